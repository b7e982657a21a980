"""interpolate_and_commit for ONE column spread over the GPUs of one box (a single oversized
column; BASELINE config C3 on 2/4/8 GPUs).

Columns cannot be dealt out when there is only one, so the transform itself is shared: four-step NTT with one all-to-all.
A size-N transform is a 2^l1 x (N / 2^l1) matrix problem (first digit k1 x inner position i):

  phase 0    every GPU runs the first pass (size-2^l1 transforms + the inter-pass twiddles) on ITS inner positions
             [g * inner / G, ..) -- the host uploads exactly that slab, so nothing is communicated before it;
  exchange   one all-to-all turns "all k1, my inner positions" into "my k1 rows, all inner positions";
  phase 1    every GPU runs the remaining passes on its rows; its outputs are the natural indices k = k1 + 2^l1 * q,
             i.e. runs of 2^l1 / G consecutive values.

interpolate_and_commit (src/starks/prover.rs:126-159) is an inverse transform, `blowup` coset transforms
(evaluate_offset_fft, prover.rs:106-123) and a tree, so one column costs three redistributions besides the two
transposes: coefficients "runs" -> slabs of the forward transform, and LDE "runs" -> contiguous row blocks for the
row-block tree of distributed.py (each rank hashes its rows and builds that subtree; subtree roots are gathered).
Every redistribution is the same routine: both sides derive, from index arithmetic alone, which elements go where
(`_plan`), pack them in increasing index order, and one all_to_all_single moves them.  Plans are cached per shape.

torch.distributed is the plumbing (NCCL over NVLink; gloo in the CPU tests where the transform phases come from a CPU
double); the transform phases are the library's kernels (s252_ntt_shared).
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _native as N
from . import distributed as D
from . import fri_distributed as F

_PLANS = {}


def _plan(key, total, world, rank, device, owner_src, owner_dst):
    """Who sends what: owner_src(idx) / owner_dst(idx) map a tensor of global indices to the rank that holds the element
    before / needs it after.  Returns (send_idx, send_counts, recv_idx, recv_counts): elements travel in increasing index
    order per (source, destination) pair, so both sides agree without exchanging any index."""
    key = key + (total, world, rank, str(device))
    if key in _PLANS:
        return _PLANS[key]
    idx = torch.arange(total, dtype=torch.int64, device=device)
    src, dst = owner_src(idx), owner_dst(idx)
    mine = idx[src == rank]
    order = torch.sort(dst[mine], stable=True).indices
    send_idx = mine[order]
    send_counts = torch.bincount(dst[mine], minlength=world).tolist()
    need = idx[dst == rank]
    order = torch.sort(src[need], stable=True).indices
    recv_idx = need[order]
    recv_counts = torch.bincount(src[need], minlength=world).tolist()
    _PLANS[key] = (send_idx, send_counts, recv_idx, recv_counts)
    return _PLANS[key]


def _redistribute(src_buf, dst_buf, plan, group):
    """dst_buf[recv_idx] = src_buf[send_idx] across the ranks (buffers: [total, 4] elements)."""
    send_idx, send_counts, recv_idx, recv_counts = plan
    send = src_buf.index_select(0, send_idx)
    recv = torch.empty((recv_idx.shape[0], 4), dtype=src_buf.dtype, device=src_buf.device)
    if dist.get_world_size(group) > 1:
        dist.all_to_all_single(recv, send, output_split_sizes=recv_counts, input_split_sizes=send_counts, group=group)
    else:
        recv = send
    dst_buf.index_copy_(0, recv_idx, recv)


class GpuColumnBackend(D.GpuBackend):
    """The transform phases on this rank's GPU (s252_ntt_shared) + what distributed.py / fri_distributed.py need."""

    def geometry(self, log_n):
        l1 = C.c_uint()
        self.ctx.check(N.lib().s252_ntt_shared(self.ctx.handle, log_n, 1, 1, 0, 2, 0, 1, None, None, None, C.byref(l1)), N.FFTError)
        return int(l1.value)

    def new_tensor(self, shape):
        return torch.empty(shape, dtype=torch.int64, device=self.device)

    def load_slab(self, host_column, buf, l1_rows, inner, lo, hi):
        """host column (LW, pinned tensor or array [N, 4]) -> buf[pos * inner + i] for i in [lo, hi), library-internal format."""
        h = host_column if torch.is_tensor(host_column) else torch.from_numpy(np.ascontiguousarray(host_column).view(np.int64))
        self._keep = h                                     # the DMA reads it asynchronously
        self.ctx.check(N.lib().s252_copy_2d_to_device(self.ctx.handle, C.c_void_p(buf.data_ptr() + 32 * lo), 32 * inner,
                                                      C.c_void_p(h.data_ptr() + 32 * lo), 32 * inner, 32 * (hi - lo), l1_rows))
        self.ctx.check(N.lib().s252_convert_elements(self.ctx.handle, C.c_void_p(buf.data_ptr()), C.c_void_p(buf.data_ptr()), buf.shape[0], 1))

    def ntt_shared(self, log_n, inverse, n_cosets, coset_offset, phase, part, parts, src, z, out):
        self.ctx.check(N.lib().s252_ntt_shared(self.ctx.handle, log_n, int(inverse), n_cosets, coset_offset, phase, part, parts,
                                               C.c_void_p(src.data_ptr()), C.c_void_p(z.data_ptr()), C.c_void_p(out.data_ptr()), None), N.FFTError)

    def read_elements(self, t):
        """device elements (internal format) -> uint64[n, 4] LW on the host."""
        tmp = torch.empty_like(t)
        self.ctx.check(N.lib().s252_convert_elements(self.ctx.handle, C.c_void_p(t.data_ptr()), C.c_void_p(tmp.data_ptr()), t.shape[0], 0))
        self.ctx.synchronize()
        return tmp.cpu().numpy().view(np.uint64)


class ShardedColumn:
    """What one rank holds after the sharded commit of one column."""

    def __init__(self, commit, coeff_runs, l1, log_n, world):
        self.commit = commit                # ShardedCommit over the LDE rows (row blocks)
        self.root = commit.root
        self.coeff_runs = coeff_runs        # [N, 4] buffer: this rank's coefficients k with (k mod 2^l1) in its run range are valid
        self.l1, self.log_n, self.world = l1, log_n, world

    def open(self, indices):
        return self.commit.open(indices)

    def free(self):
        self.commit.free()
        self.coeff_runs = None


def interpolate_and_commit_column_sharded(host_column, log_n, blowup, coset_offset, transcript, be, group=None):
    """interpolate_and_commit (prover.rs:126-159) of a one-column trace of 2^log_n rows (host_column: LW elements, the same
    array on every rank -- each rank reads only its slab).  Appends the root to the transcript; returns a ShardedColumn."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if world & (world - 1):
        raise ValueError("the number of ranks must be a power of two")
    n = 1 << log_n
    m = n * blowup
    with D.backend_scope(be):
        l1 = be.geometry(log_n)
        rows1, inner = 1 << l1, n >> l1
        if rows1 % world or inner % world or (m // world) & (m // world - 1):
            raise ValueError("too many ranks for a column of 2^%d rows" % log_n)
        w_rows, w_inner = rows1 // world, inner // world
        # ownership predicates over flat indices
        own_slab = lambda idx: (idx % inner) // w_inner                         # noqa: E731  z / input: index = (c*N +) k1*inner + i
        own_rows = lambda idx: ((idx % n) // inner) // w_rows                   # noqa: E731  z after the transpose: rows k1
        own_runs = lambda idx: (idx % rows1) // w_rows                          # noqa: E731  natural order k = k1 + rows1*q: runs
        own_lde_runs = lambda idx: ((idx // blowup) % rows1) // w_rows          # noqa: E731  LDE index j = k*blowup + c
        own_block = lambda idx: idx // (m // world)                             # noqa: E731  contiguous row blocks
        dev = be.device
        # ---- compute_trace_polys: interpolate_fft (trace.rs:104-110)
        col = be.new_tensor((n, 4))
        be.load_slab(host_column, col, rows1, inner, rank * w_inner, (rank + 1) * w_inner)
        z = be.new_tensor((n, 4))
        coeffs = be.new_tensor((n, 4))
        be.ntt_shared(log_n, True, 1, 0, 0, rank, world, col, z, coeffs)
        _redistribute(z, z, _plan(("transpose", n, rows1, 1), n, world, rank, dev, own_slab, own_rows), group)
        be.ntt_shared(log_n, True, 1, 0, 1, rank, world, col, z, coeffs)
        # ---- compute_lde_trace_evaluations: evaluate_offset_fft(blowup, Some(N), h) (prover.rs:106-123)
        # coefficients: runs -> the slabs the forward transform's first pass reads
        _redistribute(coeffs, col, _plan(("runs2slab", n, rows1), n, world, rank, dev, own_runs, own_slab), group)
        zc = be.new_tensor((m, 4))
        lde = be.new_tensor((m, 4))
        be.ntt_shared(log_n, False, blowup, coset_offset, 0, rank, world, col, zc, lde)
        _redistribute(zc, zc, _plan(("transpose", n, rows1, blowup), m, world, rank, dev, own_slab, own_rows), group)
        be.ntt_shared(log_n, False, blowup, coset_offset, 1, rank, world, col, zc, lde)
        del zc, z, col
        # ---- batch_commit (prover.rs:96-104): LDE runs -> contiguous row blocks, row-block tree
        block = be.new_tensor((m, 4))
        _redistribute(lde, block, _plan(("runs2block", n, rows1, blowup), m, world, rank, dev, own_lde_runs, own_block), group)
        rows_per = m // world
        mine = block[rank * rows_per:(rank + 1) * rows_per].clone()
        del block, lde
        sc = F.commit_row_block(mine.view(1, rows_per, 4), m, transcript, be, group)
        return ShardedColumn(sc, coeffs, l1, log_n, world)
