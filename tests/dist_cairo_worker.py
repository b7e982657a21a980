"""Worker for the sharded Cairo prover (launched under torch.distributed.run, one GPU per rank):
generate_cairo_proof_sharded must return the single-GPU prover's bytes."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import lambdaworks_cairo_prover_b200 as P                                    # noqa: E402
from lambdaworks_cairo_prover_b200 import cairo                                # noqa: E402
from lambdaworks_cairo_prover_b200.cairo_distributed import generate_cairo_proof_sharded   # noqa: E402


def main_gloo(fib_n, opts):
    """CPU ranks: the orchestration with the oracle double of the GPU backend == the oracle's own prover."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from cairo_oracle_backend import OracleCairoBackend
    from oracle.cairo_prover import cairo_prove
    dist.init_process_group("gloo")
    rank = dist.get_rank()
    regs, mem, size = cairo.run_program(cairo.fibonacci_program(fib_n))
    trace = cairo.build_main_trace(regs, mem, size)
    table = np.array(trace.table).reshape(trace.n_rows(), trace.n_cols, 4)
    # fri_collapse_log 2: FRI layers stay sharded (row-block trees, pairwise fold exchange) down to 8 evaluations;
    # None: the default threshold, far above these sizes, so the whole commit phase runs on rank 0 after one gather
    # the last case pipelines the round-1 commits in two column groups per rank (what traces of 2^20 rows and more do)
    for exchange, collapse, groups in (("a2a", 2, 1), ("p2p", None, 1), ("p2p", 4, 2)):
        proof = generate_cairo_proof_sharded(trace, opts, OracleCairoBackend(table, trace.pub_inputs), exchange=exchange,
                                             fri_collapse_log=collapse, pipeline_groups=groups)
        if rank == 0:
            want = cairo_prove(table, trace.pub_inputs, opts, threads=1).serialize()
            assert proof == want, "sharded proof differs from the oracle's (%d vs %d bytes)" % (len(proof), len(want))
        else:
            assert proof is None
    dist.barrier()
    if rank == 0:
        print("DIST_CAIRO_OK", dist.get_world_size(), trace.n_rows(), len(proof))
    dist.destroy_process_group()


def main():
    fib_n = int(sys.argv[1])
    opts = P.ProofOptions(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]))
    if len(sys.argv) > 6 and sys.argv[6] == "gloo":
        return main_gloo(fib_n, opts)
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank = dist.get_rank()
    ctx = P.Context(local)
    regs, mem, size = cairo.run_program(cairo.fibonacci_program(fib_n))
    trace = cairo.build_main_trace(regs, mem, size)
    want = cairo.generate_cairo_proof(trace, opts, ctx) if rank == 0 else None
    for collapse, groups in ((None, None), (5, 2), (9, 1)):
        proof = generate_cairo_proof_sharded(trace, opts, ctx, fri_collapse_log=collapse, pipeline_groups=groups)
        if rank == 0:
            assert proof == want, "sharded proof differs from the single-GPU proof (%d vs %d bytes)" % (len(proof), len(want))
        else:
            assert proof is None
    dist.barrier()
    if rank == 0:
        print("DIST_CAIRO_OK", dist.get_world_size(), trace.n_rows(), len(proof))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
