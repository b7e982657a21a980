"""Times generate_cairo_proof (GPU) on the regenerated fib(1,1,n) trace and prints the per-kernel
breakdown.  Usage: python tools/cairo_prove_timing.py [n=70000] [reps=5] [out.json]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import lambdaworks_cairo_prover_b200 as P  # noqa: E402
from lambdaworks_cairo_prover_b200 import cairo  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 70000
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    out_path = sys.argv[3] if len(sys.argv) > 3 else None
    t0 = time.perf_counter()
    regs, mem, size = cairo.run_program(cairo.fibonacci_program(n))
    t1 = time.perf_counter()
    trace = cairo.build_main_trace(regs, mem, size)
    t2 = time.perf_counter()
    ctx = P.Context(0)
    res = {"program": "fibonacci_%d" % n, "trace_rows": trace.n_rows(), "vm_s": t1 - t0, "build_main_trace_s": t2 - t1, "runs": {}}
    for name, opts in (("default_test_options", P.ProofOptions.default_test_options()),
                       ("Provable80Bits", P.ProofOptions.new_secure("Provable80Bits", 3))):
        for _ in range(2):
            proof = cairo.generate_cairo_proof(trace, opts, ctx)
        ctx.synchronize()
        times = []
        for _ in range(reps):
            a = time.perf_counter()
            proof = cairo.generate_cairo_proof(trace, opts, ctx)
            ctx.synchronize()
            times.append((time.perf_counter() - a) * 1e3)
        from lambdaworks_cairo_prover_b200 import _native as N
        stages = json.loads(N.lib().s252_cairo_last_prove_stages().decode())
        ctx.profile(True, reset=True)
        cairo.generate_cairo_proof(trace, opts, ctx)
        prof = ctx.profile_read()
        ctx.profile(False)
        kern = {k: round(v["ms"], 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}
        res["runs"][name] = {"ms": [round(t, 2) for t in times], "ms_median": sorted(times)[len(times) // 2], "proof_bytes": len(proof), "stages_ms": stages,
                             "kernel_ms": kern, "kernel_ms_total": round(sum(kern.values()), 2)}
    print(json.dumps(res, indent=1))
    if out_path:
        json.dump(res, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
