"""The C++ host layer (include/stark252_b200.hpp): compiles and links on CPU; on the GPU its
results are compared with the oracle."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from lambdaworks_cairo_prover_b200 import _native as N
from oracle import pyoracle as O

SRC = os.path.join(ROOT, "tests", "cpp", "host_api_demo.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "host_api_demo")


def build_demo():
    N.lib()
    libdir = os.path.dirname(N.library_path())
    deps = [SRC, N.library_path()] + [os.path.join(ROOT, "include", h) for h in ("stark252_b200.hpp", "stark252_cairo.hpp")]
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), SRC, "-o", EXE,
                               "-L", libdir, "-lstark252_b200", "-Wl,-rpath," + libdir])
    return EXE


def test_cpp_layer_compiles_and_links():
    exe = build_demo()
    out = subprocess.run([exe, "--link-only"], capture_output=True, text=True, check=True).stdout
    assert out.strip() == "len 16"


def test_cpp_cairo_front_end():
    """include/stark252_cairo.hpp: run_program + build_main_trace of the reference's `mul` program (host only)."""
    exe = build_demo()
    out = subprocess.run([exe, "--cairo-front-end"], capture_output=True, text=True, check=True).stdout
    assert out.strip() == "steps 3 rows 8 cols 34 ap_final 9 rc 32766 32769"


@pytest.mark.gpu
def test_cpp_cairo_proof_matches_oracle():
    from lambdaworks_cairo_prover_b200 import ProofOptions, cairo
    from oracle.cairo_prover import cairo_prove
    exe = build_demo()
    res = subprocess.run([exe, "--cairo-front-end", "--prove"], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    regs, mem, size = cairo.run_program([0x480680017fff8000, 6, 0x400680017fff7fff, 6, 0x208b7fff7fff7ffe])
    t = cairo.build_main_trace(regs, mem, size)
    table = np.array(t.table).reshape(t.n_rows(), t.n_cols, 4)
    want = cairo_prove(table, t.pub_inputs, ProofOptions.default_test_options(), threads=1).serialize()
    assert res.stdout.strip().splitlines()[-1] == "proof %d %s" % (len(want), O.keccak256(want).hex())


@pytest.mark.gpu
def test_cpp_layer_matches_oracle():
    exe = build_demo()
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    got = dict(line.split(" ", 1) for line in res.stdout.strip().splitlines())
    n, c, blowup = 16, 3, 4
    M64 = 2**64 - 1
    trace = np.zeros((n * c, 4), dtype=np.uint64)
    for i in range(n * c):
        trace[i, 0] = ((0x0123456789abcdef ^ ((i * 0x9e3779b97f4a7c15 & M64) >> 8))) & ((1 << 59) - 1)
        trace[i, 1] = (i * 0xbf58476d1ce4e5b9) & M64
        trace[i, 2] = (~i) & M64
        trace[i, 3] = i + 7
    want = O.interpolate_and_commit(trace.reshape(n, c, 4), blowup, 3)
    assert got["root"] == want["root"].hex()
    assert got["lde_1_5"] == "".join("%016x" % int(x) for x in want["lde"][1][5])
    assert got["path0"] == O.merkle_path(want["nodes"], 9)[0].tobytes().hex()
    assert got["none"] == "1" and got["fft_error"] == "1"
    t = O.Transcript()
    t.append(want["root"])
    last, roots, _, _ = O.fri_commit_phase(4, want["coeffs"][2], t, trace[0], n * blowup)
    assert got["fri_last"] == "".join("%016x" % int(x) for x in last)
    assert got["fri_root3"] == roots[3].tobytes().hex()
    assert int(got["nonce"]) == O.generate_nonce_with_grinding(t.challenge(), 9)


SHARDED_SRC = os.path.join(ROOT, "tests", "cpp", "sharded_prove_demo.cpp")
SHARDED_EXE = os.path.join(ROOT, "tests", "cpp", "sharded_prove_demo")


def build_sharded_demo():
    N.lib()
    libdir = os.path.dirname(N.library_path())
    deps = [SHARDED_SRC, N.library_path()] + [os.path.join(ROOT, "include", h) for h in ("stark252_b200.hpp", "stark252_cairo.hpp")]
    if not os.path.exists(SHARDED_EXE) or os.path.getmtime(SHARDED_EXE) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-pthread", "-I", os.path.join(ROOT, "include"), SHARDED_SRC, "-o", SHARDED_EXE,
                               "-L", libdir, "-lstark252_b200", "-Wl,-rpath," + libdir])
    return SHARDED_EXE


def test_cpp_sharded_demo_compiles_and_refuses_without_a_gpu():
    """The thread-per-GPU host of s252_cairo_prove_sharded builds against the headers; without a device it exits with 2
    (there is no CPU path)."""
    import torch
    exe = build_sharded_demo()
    if not torch.cuda.is_available():
        assert subprocess.run([exe, "1"], capture_output=True, text=True).returncode == 2


@pytest.mark.gpu
def test_cpp_sharded_prover_one_thread_per_gpu():
    """One process, one thread per GPU, NCCL inside the library: the sharded proof equals the single-GPU proof (1 rank on any box,
    2 and 4 ranks when the box has the GPUs)."""
    import torch
    exe = build_sharded_demo()
    env = dict(os.environ)
    spec = __import__("importlib.util").util.find_spec("nvidia.nccl")
    if spec and spec.submodule_search_locations:
        env.setdefault("S252_NCCL_LIB", os.path.join(list(spec.submodule_search_locations)[0], "lib", "libnccl.so.2"))
    for world in [w for w in (1, 2, 4) if w <= torch.cuda.device_count()]:
        res = subprocess.run([exe, str(world)], capture_output=True, text=True, env=env, timeout=300)
        assert res.returncode == 0, res.stdout + res.stderr
        assert "SHARDED_PROVE_DEMO_OK ranks %d" % world in res.stdout
