"""Builds lambdaworks_cairo_prover_b200/lib/libstark252_b200.so (sm_100a only, in-tree)."""
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
# S252_LIB_SUFFIX selects a kernel-variant build (experiments: tools/ntt_variants.sh); empty = the product library
SO = os.path.join(LIBDIR, "libstark252_b200%s.so" % os.environ.get("S252_LIB_SUFFIX", ""))
SOURCES = ["runtime.cu"]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    # every file the translation unit can include: the list is globbed so that it cannot drift from the sources
    deps = [f for pat in ("*.cu", "*.cuh", "*.hpp", "*.h", "*.inc") for f in glob.glob(os.path.join(CSRC, pat))]
    deps += glob.glob(os.path.join(os.path.dirname(HERE), "include", "*.h"))
    deps.append(os.path.abspath(__file__))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return SO
    os.makedirs(LIBDIR, exist_ok=True)
    cmd = [
        nvcc_path(), "-std=c++17", "-O3", "-lineinfo",
        "-gencode", "arch=compute_100a,code=sm_100a",
        "-shared", "-Xcompiler", "-fPIC", "-cudart", "static",
        "-ccbin", shutil.which("g++") or "g++",
        "-o", SO,
    ] + os.environ.get("S252_NVCC_FLAGS", "").split() + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(SO)
