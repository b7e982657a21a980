"""Worker for s252_cairo_prove_sharded (the whole sharded prover inside the library, NCCL called from C++): one plain python
process per GPU, the communicator id travels through a file; rank 0 compares the bytes with the single-GPU prover's.
usage: dist_native_cairo_worker.py rank world fib_n blowup queries offset grinding idfile"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import lambdaworks_cairo_prover_b200 as P                     # noqa: E402
from lambdaworks_cairo_prover_b200 import cairo                 # noqa: E402
from lambdaworks_cairo_prover_b200 import sharded as S         # noqa: E402


def main():
    rank, world, fib_n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    opts = P.ProofOptions(int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6]), int(sys.argv[7]))
    idfile = sys.argv[8]
    ctx = P.Context(rank)
    if rank == 0:
        with open(idfile + ".tmp", "wb") as f:
            f.write(S.unique_id())
        os.rename(idfile + ".tmp", idfile)
    t0 = time.time()
    while not os.path.exists(idfile):
        if time.time() - t0 > 120:
            raise RuntimeError("no communicator id after 120 s")
        time.sleep(0.05)
    comm = S.Communicator(ctx, open(idfile, "rb").read(), rank, world)
    regs, mem, size = cairo.run_program(cairo.fibonacci_program(fib_n))
    trace = cairo.build_main_trace(regs, mem, size)
    want = cairo.generate_cairo_proof(trace, opts, ctx) if rank == 0 else None
    # default threshold (the whole FRI phase on rank 0 after one gather), then FRI layers kept sharded down to 32 / 512 evaluations
    # with two and three column groups per rank in the round-1 commits
    for collapse, groups in ((None, 0), ("5", 2), ("9", 3)):
        if collapse is None:
            os.environ.pop("S252_FRI_COLLAPSE_LOG", None)
        else:
            os.environ["S252_FRI_COLLAPSE_LOG"] = collapse
        for _ in range(2):
            proof = S.generate_cairo_proof_sharded(trace, opts, comm, pipeline_groups=groups)
            if rank == 0:
                assert proof == want, "sharded proof differs from the single-GPU proof (%d vs %d bytes), collapse %s groups %d" % (
                    len(proof), len(want), collapse, groups)
            else:
                assert proof is None
    # a trace that violates the AIR: H exceeds its degree bound on every rank alike -> the error comes back on EVERY rank (nobody
    # is left waiting in a collective), and the communicator stays usable
    import numpy as np
    os.environ.pop("S252_FRI_COLLAPSE_LOG", None)
    bad = np.array(trace.table).reshape(trace.n_rows(), trace.n_cols, 4).copy()
    bad[5, 24] = bad[6, 24]
    bad[5, 24, 3] ^= 12345
    tb = cairo.trace_from_table(bad.reshape(-1, 4), trace.n_cols, trace.pub_inputs)
    try:
        S.generate_cairo_proof_sharded(tb, opts, comm)
        raise AssertionError("an unsatisfied AIR must be refused")
    except P.Stark252Error:
        pass
    proof = S.generate_cairo_proof_sharded(trace, opts, comm)
    assert (proof == want) if rank == 0 else proof is None
    comm.close()
    ctx.close()
    print("NATIVE_CAIRO_OK rank %d of %d rows %d" % (rank, world, trace.n_rows()))


if __name__ == "__main__":
    main()
