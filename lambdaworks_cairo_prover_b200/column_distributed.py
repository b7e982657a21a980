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


def _a2a_blocks(send, group):
    """send: [G, ...] (slot d goes to rank d) -> recv: [G, ...] (slot s came from rank s); equal splits."""
    if dist.get_world_size(group) == 1:
        return send
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv.view(-1), send.contiguous().view(-1), group=group)
    return recv


# The four redistributions as strided views: the ownership patterns are blocks of a 4- or 5-dimensional view of the
# flat buffer, so packing and unpacking are two tensor copies at HBM speed and the all-to-all has equal splits.  (The
# index-plan routine above stays as the general fallback and as the specification the tests were first written for.)
def _transpose_z(z, cosets, rows1, inner, world, rank, group):
    """[c][k1][i]: "all k1, my i range" -> "my k1 range, all i"."""
    if world == 1:
        return
    wr, wi = rows1 // world, inner // world
    v = z.view(cosets, world, wr, world, wi, 4)                       # [c, k1 block, k1, i block, i, limb]
    recv = _a2a_blocks(v[:, :, :, rank].permute(1, 0, 2, 3, 4).contiguous(), group)       # [d][c, wr, wi, 4]
    v[:, rank] = recv.permute(1, 2, 0, 3, 4)                                              # [c, wr, s, wi, 4]


def _runs_to_slabs(coeffs, col, rows1, inner, world, rank, group):
    """natural index k = k1 + rows1*q.  Own: k1 in my range.  Need: (k mod inner) in my range of width inner/world.
    With inner = rows1 * R and world | R this is q_lo in my range of width R/world (q = q_lo + R*q_hi)."""
    if world == 1:
        col.copy_(coeffs)
        return
    r_ = inner // rows1
    wr = rows1 // world
    src = coeffs.view(-1, world, r_ // world, world, wr, 4)            # [q_hi, q_lo block, q_lo, k1 block, k1, limb]
    dst = col.view(-1, world, r_ // world, world, wr, 4)
    recv = _a2a_blocks(src[:, :, :, rank].permute(1, 0, 2, 3, 4).contiguous(), group)     # [d][q_hi, R/G, wr, 4]
    dst[:, rank] = recv.permute(1, 2, 0, 3, 4)                                            # [q_hi, R/G, s, wr, 4]


def _runs_to_blocks(lde, block_mine, rows1, blowup, world, rank, group):
    """LDE index j = (k1 + rows1*q)*blowup + c.  Own: k1 in my range.  Need: the contiguous block j in [rank*m/G, ..), i.e.
    q in my range of width Q/world.  block_mine: [m/world, 4] receives this rank's rows."""
    if world == 1:
        block_mine.copy_(lde)
        return
    wr = rows1 // world
    src = lde.view(world, -1, world, wr, blowup, 4)                    # [q block, q, k1 block, k1, c, limb]
    recv = _a2a_blocks(src[:, :, rank].contiguous(), group)            # [d][Q/G, wr, b, 4]
    block_mine.view(-1, world, wr, blowup, 4).copy_(recv.permute(1, 0, 2, 3, 4))          # [Q/G, s, wr, b, 4]


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


def interpolate_and_commit_column_sharded(host_column, log_n, blowup, coset_offset, transcript, be, group=None, timings=None,
                                          force_index_plans=False):
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
        import time
        clock = [time.perf_counter()]

        def mark(name):
            if timings is not None:
                if hasattr(be, "sync"):
                    be.sync()
                now = time.perf_counter()
                timings[name] = timings.get(name, 0.0) + (now - clock[0]) * 1e3
                clock[0] = now
        structured = not force_index_plans and (inner // rows1) % world == 0 and inner >= rows1
        # ---- compute_trace_polys: interpolate_fft (trace.rs:104-110)
        col = be.new_tensor((n, 4))
        be.load_slab(host_column, col, rows1, inner, rank * w_inner, (rank + 1) * w_inner)
        z = be.new_tensor((n, 4))
        coeffs = be.new_tensor((n, 4))
        mark("upload")
        be.ntt_shared(log_n, True, 1, 0, 0, rank, world, col, z, coeffs)
        mark("transform")
        if structured:
            _transpose_z(z, 1, rows1, inner, world, rank, group)
        else:
            _redistribute(z, z, _plan(("transpose", n, rows1, 1), n, world, rank, dev, own_slab, own_rows), group)
        mark("exchange")
        be.ntt_shared(log_n, True, 1, 0, 1, rank, world, col, z, coeffs)
        mark("transform")
        # ---- compute_lde_trace_evaluations: evaluate_offset_fft(blowup, Some(N), h) (prover.rs:106-123)
        # coefficients: runs -> the slabs the forward transform's first pass reads
        if structured:
            _runs_to_slabs(coeffs, col, rows1, inner, world, rank, group)
        else:
            _redistribute(coeffs, col, _plan(("runs2slab", n, rows1), n, world, rank, dev, own_runs, own_slab), group)
        mark("exchange")
        zc = be.new_tensor((m, 4))
        lde = be.new_tensor((m, 4))
        be.ntt_shared(log_n, False, blowup, coset_offset, 0, rank, world, col, zc, lde)
        mark("transform")
        if structured:
            _transpose_z(zc, blowup, rows1, inner, world, rank, group)
        else:
            _redistribute(zc, zc, _plan(("transpose", n, rows1, blowup), m, world, rank, dev, own_slab, own_rows), group)
        mark("exchange")
        be.ntt_shared(log_n, False, blowup, coset_offset, 1, rank, world, col, zc, lde)
        mark("transform")
        del zc, z, col
        # ---- batch_commit (prover.rs:96-104): LDE runs -> contiguous row blocks, row-block tree
        rows_per = m // world
        mine = be.new_tensor((rows_per, 4))
        if structured:
            _runs_to_blocks(lde, mine, rows1, blowup, world, rank, group)
        else:
            block = be.new_tensor((m, 4))
            _redistribute(lde, block, _plan(("runs2block", n, rows1, blowup), m, world, rank, dev, own_lde_runs, own_block), group)
            mine.copy_(block[rank * rows_per:(rank + 1) * rows_per])
            del block
        del lde
        mark("exchange")
        sc = F.commit_row_block(mine.view(1, rows_per, 4), m, transcript, be, group)
        mark("hash")
        return ShardedColumn(sc, coeffs, l1, log_n, world)
