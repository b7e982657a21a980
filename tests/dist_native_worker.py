"""Worker for the C-ABI sharded commit (s252_interpolate_and_commit_sharded: NCCL called from the library, no torch.distributed):
one plain python process per GPU, the communicator id travels through a file.  Root, opened rows and authentication paths are
compared with the oracle's single-table interpolate_and_commit on every rank.
usage: dist_native_worker.py rank world logn n_cols blowup groups idfile [host|device]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import lambdaworks_cairo_prover_b200 as P                     # noqa: E402
from lambdaworks_cairo_prover_b200 import sharded as S         # noqa: E402
from oracle import pyoracle as O                               # noqa: E402
from util import random_felts                                  # noqa: E402


def main():
    rank, world, logn, n_cols, blowup, groups = (int(x) for x in sys.argv[1:7])
    idfile = sys.argv[7]
    mem = sys.argv[8] if len(sys.argv) > 8 else "host"
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
    uid = open(idfile, "rb").read()
    comm = S.Communicator(ctx, uid, rank, world)
    n = 1 << logn
    trace = random_felts(4242, n * n_cols).reshape(n, n_cols, 4)          # the same table on every rank
    lo, hi = S.my_columns(n_cols, world, rank)
    tables = []
    for g in range(min(groups, hi - lo)):
        a, b = S.my_columns(hi - lo, min(groups, hi - lo), g)
        tables.append(np.ascontiguousarray(trace[:, lo + a:lo + b]))
    for rep in range(2):                                                    # twice: buffers of the first commit are recycled
        if mem == "device":
            import torch
            dev = [torch.from_numpy(t.view(np.int64)).to("cuda:%d" % rank) for t in tables]
            torch.cuda.synchronize()
            sc = S.interpolate_and_commit_sharded(None, n, n_cols, blowup, 3, comm, device_pointers=[(d.data_ptr(), d.shape[1]) for d in dev])
        else:
            sc = S.interpolate_and_commit_sharded(tables, n, n_cols, blowup, 3, comm)
        want = O.interpolate_and_commit(trace, blowup, 3, want_nodes=True)
        assert sc.root == bytes(want["root"]), "root differs from the oracle's"
        m = n * blowup
        idx = sorted({0, 1, m // 2 - 1, m // 2, m - 1, (m // world) * (world - 1), 7 % m, (3 * m) // 4 + 1})
        rows, paths = sc.open(idx)
        for k, i in enumerate(idx):
            assert (rows[k] == want["lde"][:, i]).all(), ("row", i)
            assert (paths[k] == np.asarray(O.merkle_path(want["nodes"], i))).all(), ("path", i)
        try:
            sc.open([m])
            raise AssertionError("an out-of-range position must be refused")
        except P.Stark252Error as e:
            assert e.code == -4
        sc.free()
    comm.close()
    ctx.close()
    print("NATIVE_SHARDED_OK rank %d of %d" % (rank, world))


if __name__ == "__main__":
    main()
