"""ONE Cairo proof (fib(1,1,n), Provable80Bits options) sharded over the GPUs of the box: wall clock per
proof, barrier to barrier, max over ranks.  Run under torchrun; rank 0 prints one JSON line.
Usage: torchrun ... tools/cairo_sharded_timing.py [n=70000] [reps=5]"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import lambdaworks_cairo_prover_b200 as P                                    # noqa: E402
from lambdaworks_cairo_prover_b200 import cairo                                # noqa: E402
from lambdaworks_cairo_prover_b200.cairo_distributed import generate_cairo_proof_sharded   # noqa: E402


def main():
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    fib_n = int(sys.argv[1]) if len(sys.argv) > 1 else 70000
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    exchange = sys.argv[3] if len(sys.argv) > 3 else "a2a"
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    ctx = P.Context(local)
    regs, mem, size = cairo.run_program(cairo.fibonacci_program(fib_n))
    trace = cairo.build_main_trace(regs, mem, size)
    golden_ok = None
    golden = os.path.join(ROOT, "tests", "golden", "reference_proofs", "fibonacci_%d.proof" % fib_n)
    proof = generate_cairo_proof_sharded(trace, P.ProofOptions.default_test_options(), ctx)
    if rank == 0 and fib_n == 70000 and os.path.exists(golden):
        raw = open(golden, "rb").read()
        ln = int.from_bytes(raw[:8], "big")
        golden_ok = proof == raw[8:8 + ln]
    opts = P.ProofOptions.new_secure("Provable80Bits", 3)
    for _ in range(2):
        generate_cairo_proof_sharded(trace, opts, ctx)
    times = []
    stages = {}
    for _ in range(2):
        generate_cairo_proof_sharded(trace, opts, ctx, timings=stages, exchange=exchange)
    stages = {k: (round(v / 2, 2) if not isinstance(v, dict) else {a: round(x / 2, 2) for a, x in v.items()}) for k, v in stages.items()}
    for _ in range(reps):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        proof = generate_cairo_proof_sharded(trace, opts, ctx, exchange=exchange)
        torch.cuda.synchronize()
        dist.barrier()
        times.append((time.perf_counter() - t0) * 1e3)
    tt = torch.tensor(times, device="cuda", dtype=torch.float64)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    times = sorted(tt.cpu().tolist())
    single = None
    if rank == 0:
        for _ in range(2):
            cairo.generate_cairo_proof(trace, opts, ctx)
        a = time.perf_counter()
        same = cairo.generate_cairo_proof(trace, opts, ctx) == proof
        single = (time.perf_counter() - a) * 1e3
        print(json.dumps({"metric": "cairo_fib_prove_time_one_proof_sharded", "program": "cairo0 fibonacci_%d" % fib_n, "trace_rows": trace.n_rows(),
                          "n_gpus": world, "exchange": exchange, "ms": times[len(times) // 2], "ms_all": [round(x, 2) for x in times], "proof_bytes": len(proof),
                          "same_bytes_as_single_gpu": same, "single_gpu_ms_same_process": round(single, 2),
                          "golden_proof_byte_identical": golden_ok, "stages_ms_rank0_synchronised": stages,
                          "timing": "barrier-to-barrier wall clock around generate_cairo_proof_sharded, max over ranks; host table in, proof bytes out"}))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
