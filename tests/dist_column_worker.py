"""Worker for the one-column sharded commit (lambdaworks_cairo_prover_b200/column_distributed.py), launched under
torch.distributed.run.

mode "gloo": CPU ranks; the two transform phases come from a python-integer restatement of the four-step split
             (TEST INFRASTRUCTURE), so the test covers the geometry, the three redistributions, the row-block tree and the
             openings against the oracle's interpolate_and_commit.
mode "nccl": one GPU per rank; the phases are the library's kernels (s252_ntt_shared)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from lambdaworks_cairo_prover_b200 import column_distributed as CD   # noqa: E402
from oracle import pyoracle as O                                      # noqa: E402
from util import random_felts                                         # noqa: E402

P = O.P


class OracleColumnBackend:
    """CPU double: tensors hold LW elements; the phases are the four-step formulas in python integers."""
    device = torch.device("cpu")

    def __init__(self, l1):
        self.l1 = l1

    def geometry(self, log_n):
        return self.l1

    def new_tensor(self, shape):
        return torch.zeros(shape, dtype=torch.int64)

    def load_slab(self, host_column, buf, l1_rows, inner, lo, hi):
        h = torch.from_numpy(np.ascontiguousarray(host_column).view(np.int64))
        buf.view(l1_rows, inner, 4)[:, lo:hi].copy_(h.view(l1_rows, inner, 4)[:, lo:hi])

    def ntt_shared(self, log_n, inverse, n_cosets, coset_offset, phase, part, parts, src, z, out):
        n = 1 << log_n
        rows1, inner = 1 << self.l1, n >> self.l1
        m = n * n_cosets
        w_n = O.lw_to_int(O.primitive_root(log_n))
        w_m = O.lw_to_int(O.primitive_root(m.bit_length() - 1))
        if inverse:
            w_n = pow(w_n, -1, P)
        za = z.numpy().view(np.uint64)
        for c in range(n_cosets):
            s = 1 if inverse else coset_offset * pow(w_m, c, P) % P
            scale = pow(n, -1, P) if inverse else 1
            if phase == 0:
                xs = [O.lw_to_int(v) for v in src.numpy().view(np.uint64)]
                w1 = pow(w_n, inner, P)                                   # order rows1
                s_i = pow(s, inner, P)
                for i in range(part * inner // parts, (part + 1) * inner // parts):
                    for k1 in range(rows1):
                        acc = sum(xs[pos * inner + i] * pow(s_i, pos, P) * pow(w1, pos * k1, P) for pos in range(rows1)) % P
                        za[c * n + k1 * inner + i] = O.int_to_lw(acc * pow(s, i, P) * pow(w_n, k1 * i, P) * scale % P)
            else:
                oa = out.numpy().view(np.uint64)
                wi = pow(w_n, rows1, P)                                   # order inner
                for k1 in range(part * rows1 // parts, (part + 1) * rows1 // parts):
                    ys = [O.lw_to_int(za[c * n + k1 * inner + i]) for i in range(inner)]
                    for q in range(inner):
                        acc = sum(ys[i] * pow(wi, i * q, P) for i in range(inner)) % P
                        oa[(k1 + rows1 * q) * n_cosets + c] = O.int_to_lw(acc)

    def commit_block(self, cols):
        arr = cols.contiguous().numpy().view(np.uint64)
        nodes, root = O.commit_columns(arr)
        return {"cols": arr, "nodes": nodes}, root

    def open_block(self, block, local_idx):
        return [block["cols"][:, i] for i in local_idx], [O.merkle_path(block["nodes"], i) for i in local_idx]

    def read_elements(self, t):
        return t.numpy().view(np.uint64)

    @staticmethod
    def keccak(data):
        return O.keccak256(data)


def main():
    mode, log_n, blowup = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    dist.init_process_group("gloo" if mode == "gloo" else "nccl")
    rank, world = dist.get_rank(), dist.get_world_size()
    n = 1 << log_n
    column = random_felts(777 + log_n, n)                                   # same on every rank
    if mode == "gloo":
        be = OracleColumnBackend(int(sys.argv[4]))
        transcript = O.Transcript()
    else:
        import lambdaworks_cairo_prover_b200 as PR
        local = int(os.environ.get("LOCAL_RANK", rank))
        torch.cuda.set_device(local)
        be = CD.GpuColumnBackend(PR.Context(local))
        transcript = PR.DefaultTranscript()
    sc = CD.interpolate_and_commit_column_sharded(column, log_n, blowup, 3, transcript, be)
    want = O.interpolate_and_commit(column.reshape(n, 1, 4), blowup, 3, threads=4)
    assert sc.root == want["root"], "rank %d: root differs" % rank
    # this rank's coefficients (runs of the natural order)
    got = be.read_elements(sc.coeff_runs)
    rows1 = 1 << sc.l1
    k = np.arange(n)
    mine = (k % rows1) // (rows1 // world) == rank
    assert (got[mine] == want["coeffs"][0][mine]).all(), "rank %d: coefficients differ" % rank
    m = n * blowup
    idx = [0, 1, m // 2 - 1, m // 2, m - 1, (m * 5) // 7]
    rows, paths = sc.open(idx)
    for q, i in enumerate(idx):
        assert (np.asarray(rows[q]).view(np.uint64).reshape(-1) == want["lde"][0, i]).all(), (rank, i)
        assert [bytes(p) for p in paths[q]] == [bytes(x) for x in O.merkle_path(want["nodes"], i)], (rank, i)
    dist.barrier()
    if rank == 0:
        print("DIST_COLUMN_OK", mode, world, log_n, sc.root.hex())
    sc.free()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
