"""Where the time of ONE sharded Cairo proof goes on this rank: synchronised stage times + the library's per-kernel CUDA-event
profile of one proof.  Run under torchrun (any world size); rank 0 prints one JSON line.
Usage: torchrun ... tools/sharded_stage_probe.py [n=280000]"""
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
    fib_n = int(sys.argv[1]) if len(sys.argv) > 1 else 280000
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    ctx = P.Context(local)
    regs, mem, size = cairo.run_program(cairo.fibonacci_program(fib_n))
    trace = cairo.build_main_trace(regs, mem, size)
    opts = P.ProofOptions.new_secure("Provable80Bits", 3)
    for _ in range(3):
        generate_cairo_proof_sharded(trace, opts, ctx)
    times = []
    for _ in range(3):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        generate_cairo_proof_sharded(trace, opts, ctx)
        torch.cuda.synchronize()
        dist.barrier()
        times.append((time.perf_counter() - t0) * 1e3)
    stages = {}
    generate_cairo_proof_sharded(trace, opts, ctx, timings=stages)
    ctx.profile(True, reset=True)
    generate_cairo_proof_sharded(trace, opts, ctx)
    ctx.synchronize()
    prof = ctx.profile_read()
    ctx.profile(False)
    single = None
    if world == 1:
        for _ in range(2):
            cairo.generate_cairo_proof(trace, opts, ctx)
        a = time.perf_counter()
        cairo.generate_cairo_proof(trace, opts, ctx)
        single = (time.perf_counter() - a) * 1e3
    if rank == 0:
        kern = {k: {"ms": round(v["ms"], 3), "launches": v["launches"]} for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}
        print(json.dumps({"program": "cairo0 fibonacci_%d" % fib_n, "n_gpus": world, "ms_all": [round(x, 2) for x in times], "stages_ms": stages,
                          "kernel_ms_total": round(sum(v["ms"] for v in prof.values()), 2), "kernels": kern, "single_gpu_native_ms": single}))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
