"""Wall-clock breakdown of one bench step (host + device), to find non-kernel time."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import lambdaworks_cairo_prover_b200 as P
from lambdaworks_cairo_prover_b200 import _native as N, felt

ctx = P.Context(0)
L = N.lib()
n = 1 << 19
m = 4 * n
bufs = {}
for k, (seed, cnt) in {"main": (1, n * 34), "aux": (2, n * 18), "comp": (3, n * 2), "p0": (4, n)}.items():
    a = bench.splitmix_felts(seed, cnt)
    p = ctx.device_alloc(a.nbytes)
    ctx.to_device(p, a)
    bufs[k] = p
off = felt.from_int(3)
root = np.empty(32, dtype=np.uint8)


def T(label, fn):
    ctx.synchronize()
    t0 = time.perf_counter()
    r = fn()
    ctx.synchronize()
    print("  %-28s %8.3f ms" % (label, (time.perf_counter() - t0) * 1e3))
    return r


for it in range(3):
    print("iteration", it)
    ctx.profile(True, reset=True)
    tr = P.DefaultTranscript()
    hs = []
    def ic(key, cols):
        h = C.c_void_p()
        ctx.check(L.s252_interpolate_and_commit(ctx.handle, C.c_void_p(bufs[key]), n, cols, 4, 3, N.DEVICE, C.byref(h), N.ptr(root)))
        hs.append(h)
    T("interpolate_and_commit main", lambda: ic("main", 34))
    T("interpolate_and_commit aux", lambda: ic("aux", 18))
    def lc():
        h = C.c_void_p()
        ctx.check(L.s252_lde_and_commit(ctx.handle, C.c_void_p(bufs["comp"]), n, 2, n, 4, 3, N.DEVICE, C.byref(h), N.ptr(root)))
        hs.append(h)
    T("lde_and_commit comp", lc)
    fh = C.c_void_p()
    last = np.empty(4, dtype=np.uint64)
    roots = np.empty((19, 32), dtype=np.uint8)
    T("fri_commit_phase", lambda: ctx.check(L.s252_fri_commit_phase(ctx.handle, 19, C.c_void_p(bufs["p0"]), n, tr.handle, N.ptr(off), m, N.DEVICE, C.byref(fh), N.ptr(last), N.ptr(roots))))
    ch = tr.challenge()
    T("grinding", lambda: P.generate_nonce_with_grinding(ch, 20, ctx))
    def fr():
        for h in hs:
            L.s252_commit_destroy(h)
        L.s252_fri_destroy(fh)
    T("destroy", fr)
    prof = ctx.profile_read()
    for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
        print("    %-20s x%-3d %8.3f ms" % (k, v["launches"], v["ms"]))
