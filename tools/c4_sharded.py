"""C4 (BASELINE.json configs[3]): synthetic 2^22-row, 33-column trace, blowup 8, ONE commit with the
columns sharded over the GPUs of the box (run under torchrun; world size 1 = single-GPU path).
Column j of the table is its own splitmix64 stream, so every world size commits the same table
and must print the same root.  Timing: barrier-to-barrier wall clock around the commit, host shards
in pinned memory (the H2D copy of the shard is inside the timed region), best of 3."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import lambdaworks_cairo_prover_b200 as P
from lambdaworks_cairo_prover_b200 import _native as N
from lambdaworks_cairo_prover_b200 import distributed as D

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 22
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 33
blowup = int(sys.argv[3]) if len(sys.argv) > 3 else 8
groups = int(sys.argv[4]) if len(sys.argv) > 4 else 1      # pipeline groups per rank (upload/LDE of g+1 overlaps the exchange of g)
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << log_n
a, b = D.column_shards(cols, world)[rank]
shard = torch.empty((n, b - a, 4), dtype=torch.int64, pin_memory=True)
view = shard.numpy().view(np.uint64)
for j in range(a, b):
    view[:, j - a, :] = bench.splitmix_felts(0xC400 + 7919 * j, n)
tables = None
if world > 1 and groups > 1:        # TraceTable::get_cols per pipeline group, pinned (set-up, outside the timed region)
    tables = []
    for lo, hi in D.group_ranges(b - a, groups):
        t = torch.empty((n, hi - lo, 4), dtype=torch.int64, pin_memory=True)
        t.copy_(shard[:, lo:hi])
        tables.append(t)
ctx = P.Context(local)
backend = D.GpuBackend(ctx)
L = N.lib()
best, root = 1e30, None
for it in range(4):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if world == 1:
        h = C.c_void_p()
        r = np.empty(32, dtype=np.uint8)
        ctx.check(L.s252_interpolate_and_commit(ctx.handle, C.c_void_p(shard.data_ptr()), n, cols, blowup, 3, N.HOST, C.byref(h), N.ptr(r)))
        root = r.tobytes()
        L.s252_commit_destroy(h)
    else:
        sc = D.interpolate_and_commit_sharded(tables if tables is not None else view.reshape(-1, 4), n, cols, blowup, 3,
                                              P.DefaultTranscript(), backend)
        root = sc.root
        sc.free()
    ctx.synchronize()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = (time.perf_counter() - t0) * 1e3
    if it > 0:
        best = min(best, dt)
if world > 1:
    t = torch.tensor([best], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    best = float(t.item())
if rank == 0:
    elems = n * blowup * cols
    print(json.dumps({"config": "C4 2^%d x %d, blowup %d, one sharded commit" % (log_n, cols, blowup), "n_gpus": world, "pipeline_groups": groups,
                      "ms": best, "elems_per_s": elems / (best * 1e-3), "root": root.hex(),
                      "timing": "wall clock, H2D of the pinned shard included"}))
if world > 1:
    dist.destroy_process_group()
