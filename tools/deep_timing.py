"""Timing of rounds 3 and 4 on the GPU at the C2 shape (N = 2^19, 34 + 18 trace columns, blowup 4):
OOD frame (2 points x 52 polynomials + H1/H2), DEEP composition in the evaluation domain, FRI."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import lambdaworks_cairo_prover_b200 as P
from lambdaworks_cairo_prover_b200 import felt

ctx = P.Context(0)
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 19
n, blowup, h = 1 << log_n, 4, 3
dom = P.Domain(n, P.ProofOptions(blowup, 80, h, 20))
t = P.DefaultTranscript()
main, _ = P.interpolate_and_commit(P.TraceTable(bench.splitmix_felts(1, n * 34), 34), dom, t, ctx)
aux, _ = P.interpolate_and_commit(P.TraceTable(bench.splitmix_felts(2, n * 18), 18), dom, t, ctx)
comp, root = P.lde_and_commit([P.Polynomial(bench.splitmix_felts(3, n)), P.Polynomial(bench.splitmix_felts(4, n))], dom, ctx)
t.append(root)
z = P.transcript_to_field(t)
zi = felt.to_int(z)
res = {}
for rep in range(3):
    ctx.profile(True, reset=True)
    ctx.synchronize()
    t0 = time.perf_counter()
    ood = P.get_trace_evaluations([main, aux], z, [0, 1], n, ctx)
    hz = P.evaluate_at(comp, felt.from_int(zi * zi % felt.MODULUS))
    ctx.synchronize()
    t1 = time.perf_counter()
    tr = P.DefaultTranscript()
    tr.append(b"x")
    gamma, gamma_p = P.transcript_to_field(tr), P.transcript_to_field(tr)
    gammas = np.stack(P.batch_sample_challenges(2 * 52, tr))
    ctx.synchronize()
    t2 = time.perf_counter()
    last, layers = P.fri_commit_phase_deep(log_n, [main, aux], comp, z, [0, 1], ood, hz[0], hz[1], gamma, gamma_p, gammas, tr, h)
    ctx.synchronize()
    t3 = time.perf_counter()
    layers.free()
    prof = ctx.profile_read()
    res = {"log_n": log_n, "round3_ood_ms": (t1 - t0) * 1e3, "round4_deep_plus_fri_ms": (t3 - t2) * 1e3,
           "kernels_ms": {k: round(v["ms"], 3) for k, v in prof.items()}}
print(json.dumps(res))
