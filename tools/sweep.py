"""Throughput on the other BASELINE.json configs (parity-test shapes, not bench lines):
C3 single-column LDE + commit sweep, C4 (2^22 x 33, blowup 8), C5 (FRI from 2^24 + grinding 20).
Device-resident inputs, CUDA events on the library's stream, best of 3 after a warm-up call.
Writes gpurun_out/sweep_r1.json."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import lambdaworks_cairo_prover_b200 as P
from lambdaworks_cairo_prover_b200 import _native as N, felt

ctx = P.Context(0)
L = N.lib()
stream = torch.cuda.ExternalStream(ctx.stream, device=0)
results = []


def timed(fn, reps=3):
    fn()   # warm-up (tables, arena)
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx.synchronize()
        e0.record(stream)
        fn()
        e1.record(stream)
        ctx.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def upload(arr):
    p = ctx.device_alloc(arr.nbytes)
    ctx.to_device(p, arr)
    return p


def commit_case(tag, log_n, cols, blowup):
    n = 1 << log_n
    trace = bench.splitmix_felts(0xC300 + log_n + cols, n * cols)
    d = upload(trace)
    del trace
    root = np.empty(32, dtype=np.uint8)

    def run():
        h = C.c_void_p()
        ctx.check(L.s252_interpolate_and_commit(ctx.handle, C.c_void_p(d), n, cols, blowup, 3, N.DEVICE, C.byref(h), N.ptr(root)))
        L.s252_commit_destroy(h)

    ms = timed(run)
    ctx.device_free(d)
    L.s252_ctx_trim(ctx.handle)
    elems = n * blowup * cols
    r = {"config": tag, "trace_rows_log2": log_n, "columns": cols, "blowup": blowup, "ms": ms, "elems": elems,
         "elems_per_s": elems / (ms * 1e-3), "root": root.tobytes().hex()}
    results.append(r)
    print(json.dumps(r), flush=True)


def fri_case(log_n, blowup, grind):
    n = 1 << log_n
    m = n * blowup
    p0 = bench.splitmix_felts(0xC500, n)
    d = upload(p0)
    off = felt.from_int(3)
    out = {}

    def run():
        t = P.DefaultTranscript()
        t.append(bytes(32))
        fh = C.c_void_p()
        last = np.empty(4, dtype=np.uint64)
        ctx.check(L.s252_fri_commit_phase(ctx.handle, log_n, C.c_void_p(d), n, t.handle, N.ptr(off), m, N.DEVICE, C.byref(fh), N.ptr(last), None))
        out["nonce"] = P.generate_nonce_with_grinding(t.challenge(), grind, ctx)
        out["last"] = felt.to_bytes_be(last).hex()
        L.s252_fri_destroy(fh)

    ms = timed(run)
    ctx.device_free(d)
    elems = sum(m >> k for k in range(log_n))
    r = {"config": "C5 FRI commit phase from 2^%d (blowup %d) + grinding %d" % (log_n + 2, blowup, grind), "ms": ms,
         "layer_elems": elems, "elems_per_s": elems / (ms * 1e-3), **out}
    results.append(r)
    print(json.dumps(r), flush=True)


which = sys.argv[1:] or ["c3", "c4", "c5"]
if "c3" in which:
    for log_n in (16, 18, 20, 22, 24):
        for b in (4, 8):
            commit_case("C3 single column", log_n, 1, b)
    commit_case("C3 single column", 26, 1, 4)
if "c4" in which:
    commit_case("C4 2^22 x 33, blowup 8 (1 GPU)", 22, 33, 8)
if "c5" in which:
    fri_case(22, 4, 20)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(results, open("gpurun_out/sweep_r1.json", "w"), indent=1)
