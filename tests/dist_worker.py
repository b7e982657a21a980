"""Worker for tests/test_distributed.py (launched under torch.distributed.run).

mode "gloo": CPU ranks; the two compute steps are injected from the oracle, so the test covers the
sharding logic (column ranges, packing, the all-to-all, subtree roots -> top of the tree, openings).
mode "nccl": one GPU per rank; the compute steps are the library's kernels."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from lambdaworks_cairo_prover_b200 import distributed as D   # noqa: E402
from oracle import pyoracle as O                              # noqa: E402
from util import random_felts                                 # noqa: E402


class OracleBackend:
    """Test double: oracle arithmetic on CPU tensors."""

    def lde(self, shard_table, n_rows, n_cols, blowup, coset_offset):
        r = O.interpolate_and_commit(np.asarray(shard_table).reshape(n_rows, n_cols, 4), blowup, coset_offset, want_nodes=False)
        return None, torch.from_numpy(r["lde"].view(np.int64).copy())

    def commit_block(self, cols):
        arr = cols.contiguous().numpy().view(np.uint64)
        nodes, root = O.commit_columns(arr)
        return {"cols": arr, "nodes": nodes}, root

    def before_collective(self):
        pass

    def after_collective(self):
        pass

    def open_block(self, block, local_idx):
        rows = [block["cols"][:, i] for i in local_idx]
        paths = [O.merkle_path(block["nodes"], i) for i in local_idx]
        return rows, paths

    @staticmethod
    def keccak(data):
        return O.keccak256(data)


def main():
    mode, logn, n_cols, blowup = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    dist.init_process_group("gloo" if mode == "gloo" else "nccl")
    rank, world = dist.get_rank(), dist.get_world_size()
    n = 1 << logn
    trace = random_felts(4242, n * n_cols).reshape(n, n_cols, 4)          # same on every rank
    a, b = D.column_shards(n_cols, world)[rank]
    shard = np.ascontiguousarray(trace[:, a:b])                              # TraceTable::get_cols
    if mode == "gloo":
        backend = OracleBackend()
        transcript = O.Transcript()
    else:
        import lambdaworks_cairo_prover_b200 as P
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
        ctx = P.Context(int(os.environ.get("LOCAL_RANK", rank)))
        backend = D.GpuBackend(ctx)
        transcript = P.DefaultTranscript()
    groups = int(sys.argv[5]) if len(sys.argv) > 5 else 1
    exchange = sys.argv[6] if len(sys.argv) > 6 else "p2p"
    if groups > 1:
        tables = [np.ascontiguousarray(shard[:, lo:hi]) for lo, hi in D.group_ranges(b - a, groups)]
        if mode == "nccl":
            tables = [torch.from_numpy(t.view(np.int64)).pin_memory() for t in tables]
        sc = D.interpolate_and_commit_sharded(tables, n, n_cols, blowup, 3, transcript, backend)
    else:
        sc = D.interpolate_and_commit_sharded(shard.reshape(-1, 4), n, n_cols, blowup, 3, transcript, backend, exchange=exchange)
    # single-process answer
    want = O.interpolate_and_commit(trace, blowup, 3, threads=4)
    assert sc.root == want["root"], "rank %d: root differs" % rank
    t2 = O.Transcript()
    t2.append(want["root"])
    assert transcript.challenge() == t2.challenge()
    m = n * blowup
    idx = [0, 1, m // 2 - 1, m // 2, m - 1, (m * 5) // 7]
    rows, paths = sc.open(idx)
    for q, i in enumerate(idx):
        assert (np.asarray(rows[q]).view(np.uint64) == want["lde"][:, i]).all(), (rank, i)
        assert [bytes(p) for p in paths[q]] == [bytes(x) for x in O.merkle_path(want["nodes"], i)], (rank, i)
        assert O.merkle_verify(sc.root, i, np.asarray(rows[q]).view(np.uint64), paths[q])
    # several commits opened with one exchange (the sharded Cairo prover opens the main and the aux table together)
    (rows2, paths2), (rows3, paths3) = D.open_many([sc, sc], idx)
    for q in range(len(idx)):
        assert (np.asarray(rows2[q]) == np.asarray(rows[q])).all() and (np.asarray(rows3[q]) == np.asarray(rows[q])).all()
        assert [bytes(p) for p in paths2[q]] == [bytes(p) for p in paths[q]] == [bytes(p) for p in paths3[q]]
    if mode == "nccl":
        # this rank's columns: coefficients and LDE are bit-exact too
        for j in range(b - a):
            assert (sc.local.coefficients(j) == want["coeffs"][a + j]).all()
            assert (sc.local.lde_column(j) == want["lde"][a + j]).all()
    dist.barrier()
    if rank == 0:
        print("DIST_OK", mode, world, sc.root.hex())
    sc.free()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
