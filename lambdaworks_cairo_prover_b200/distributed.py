"""Column-sharded interpolate_and_commit across the GPUs of one box.

One process per GPU.  Columns are independent for interpolation and LDE, so rank g transforms its
own contiguous range of columns with no communication.  Leaves hash whole rows, so one exchange is
needed: an all-to-all that turns "my columns, all rows" into "all columns, my row block"
(rank g ends up with rows [g*M/G, (g+1)*M/G)).  Each rank hashes its block and builds that subtree;
because G is a power of two the subtree roots are exactly level log2(G) of the reference's heap,
so an all-gather of G digests and G-1 host hashes finish the tree.  The result is the same root and
the same authentication paths as the single-GPU (and the reference's) tree.

torch.distributed is the plumbing (NCCL over NVLink on GPUs; gloo in the CPU tests, where the two
compute steps are injected).  The compute steps are the library's own kernels through the C ABI.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _native as N
from .merkle import DeviceCommit


def column_shards(n_cols, world):
    """Contiguous, balanced column ranges: 33 columns over 8 ranks -> 5,4,4,4,4,4,4,4."""
    base, extra = divmod(n_cols, world)
    counts = [base + (1 if r < extra else 0) for r in range(world)]
    starts = [sum(counts[:r]) for r in range(world)]
    return [(s, s + c) for s, c in zip(starts, counts)]


def _is_pow2(n):
    return n > 0 and n & (n - 1) == 0


class _DevicePointer:
    """Zero-copy view of library-owned device memory as a torch tensor (int64 words)."""

    def __init__(self, ptr, n_words):
        self.__cuda_array_interface__ = {"shape": (n_words,), "typestr": "<i8", "data": (int(ptr), False), "version": 2}


class GpuBackend:
    """The two compute steps on this rank's GPU, through the C ABI."""

    def __init__(self, ctx):
        self.ctx = ctx
        self.device = torch.device("cuda", ctx.device)

    def lde(self, shard_table, n_rows, n_cols, blowup, coset_offset):
        """-> (handle, tensor[n_cols, M, 4] int64 view of the LDE columns)"""
        h = C.c_void_p()
        self.ctx.check(N.lib().s252_interpolate_and_lde(self.ctx.handle, N.ptr(shard_table), n_rows, n_cols, blowup,
                                                        coset_offset, N.HOST, C.byref(h)), N.FFTError)
        commit = DeviceCommit.__new__(DeviceCommit)
        commit.ctx, commit.handle, commit.root = self.ctx, h, b""
        commit.n_cols, commit.n_rows, commit.n_coeffs = n_cols, n_rows * blowup, n_rows
        self.ctx.adopt(commit)
        ptr = N.lib().s252_commit_device_lde(h)
        t = torch.as_tensor(_DevicePointer(ptr, n_cols * n_rows * blowup * 4), device=self.device)
        return commit, t.view(n_cols, n_rows * blowup, 4)

    def commit_block(self, cols):
        """cols: tensor[c_total, rows, 4] on this device -> (DeviceCommit with the subtree, subtree root bytes)"""
        c_total, rows = cols.shape[0], cols.shape[1]
        h = C.c_void_p()
        root = np.empty(32, dtype=np.uint8)
        self.ctx.check(N.lib().s252_commit_device_columns(self.ctx.handle, C.c_void_p(cols.data_ptr()), rows, c_total, rows,
                                                          C.byref(h), N.ptr(root)))
        return DeviceCommit(self.ctx, h, root.tobytes()), root.tobytes()

    def before_collective(self):
        self.ctx.synchronize()          # the library's stream -> visible to NCCL's stream

    def after_collective(self):
        torch.cuda.synchronize(self.device)

    def open_block(self, block, local_idx):
        rows, paths = block.open(local_idx)
        return rows, paths

    @staticmethod
    def keccak(data):
        out = (C.c_uint8 * 32)()
        buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
        N.lib().s252_keccak256(buf, len(data), out)
        return bytes(out)


def build_top(subtree_roots, keccak):
    """Heap (root at 0) over G subtree roots: node = Keccak256(left || right)."""
    g = len(subtree_roots)
    nodes = [b""] * (2 * g - 1)
    nodes[g - 1:] = list(subtree_roots)
    for i in range(g - 2, -1, -1):
        nodes[i] = keccak(nodes[2 * i + 1] + nodes[2 * i + 2])
    return nodes


class ShardedCommit:
    """What one rank holds after a sharded commit."""

    def __init__(self, backend, group, local, block, top, n_rows, n_cols, shards):
        self.backend, self.group = backend, group
        self.local = local          # this rank's columns: coefficients + LDE, all rows
        self.block = block          # all columns, this rank's row block + its subtree
        self.top = top              # replicated top of the tree (2G-1 digests)
        self.root = top[0]
        self.n_rows, self.n_cols, self.shards = n_rows, n_cols, shards

    def open(self, indices):
        """Rows and authentication paths (leaf -> root) for global positions, on every rank:
        open_deep_composition_poly's reads (src/starks/prover.rs:484-529)."""
        world = dist.get_world_size(self.group)
        rank = dist.get_rank(self.group)
        rows_per = self.n_rows // world
        mine = [(q, i % rows_per) for q, i in enumerate(indices) if i // rows_per == rank]
        part = {}
        if mine:
            rows, paths = self.backend.open_block(self.block, [i for _, i in mine])
            for k, (q, _) in enumerate(mine):
                part[q] = (np.asarray(rows[k]).copy(), [bytes(np.asarray(p).tobytes()) for p in paths[k]])
        gathered = [None] * world
        dist.all_gather_object(gathered, part, group=self.group)
        merged = {}
        for p in gathered:
            merged.update(p)
        out_rows, out_paths = [], []
        for q, i in enumerate(indices):
            rows, path = merged[q]
            node = (world - 1) + i // rows_per            # heap index of the owner's subtree root in `top`
            while node != 0:
                sib = node + 1 if node & 1 else node - 1
                path = path + [self.top[sib]]
                node = (node - 1) >> 1
            out_rows.append(rows)
            out_paths.append(path)
        return out_rows, out_paths

    def free(self):
        for h in (self.local, self.block):
            if h is not None and hasattr(h, "free"):
                h.free()


def interpolate_and_commit_sharded(shard_table, n_rows, n_cols_total, blowup, coset_offset, transcript, backend, group=None):
    """interpolate_and_commit (src/starks/prover.rs:126-159) for ONE trace whose columns are spread
    over the ranks of `group`.  shard_table: this rank's columns as a row-major table
    (n_rows x c_rank, i.e. TraceTable::get_cols, src/starks/trace.rs:31-43).  Every rank gets the
    root (appended to its transcript, prover.rs:151)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if not _is_pow2(world):
        raise ValueError("the number of ranks must be a power of two (subtree roots must be heap nodes)")
    shards = column_shards(n_cols_total, world)
    c_mine = shards[rank][1] - shards[rank][0]
    m = n_rows * blowup
    if m % world or not _is_pow2(m // world):
        raise ValueError("LDE rows must split evenly over the ranks")
    rows_per = m // world
    counts = [b - a for a, b in shards]
    if min(counts) == 0:
        raise ValueError("fewer columns than ranks: use fewer ranks for this table")

    local, lde = backend.lde(shard_table, n_rows, c_mine, blowup, coset_offset)      # [c_mine, M, 4]
    # pack: destination-major [G][c_mine][rows_per] so that each destination's slice is contiguous
    send = lde.view(c_mine, world, rows_per * 4).permute(1, 0, 2).contiguous().view(world * c_mine, rows_per * 4)
    recv = torch.empty((n_cols_total, rows_per * 4), dtype=send.dtype, device=send.device)
    backend.before_collective()
    dist.all_to_all_single(recv, send, output_split_sizes=counts, input_split_sizes=[c_mine] * world, group=group)
    backend.after_collective()
    block, sub_root = backend.commit_block(recv.view(n_cols_total, rows_per, 4))
    mine = torch.frombuffer(bytearray(sub_root), dtype=torch.uint8).to(send.device)
    gathered = torch.empty(32 * world, dtype=torch.uint8, device=send.device)
    dist.all_gather_into_tensor(gathered, mine, group=group)
    backend.after_collective()
    roots = bytes(gathered.cpu().numpy().tobytes())
    top = build_top([roots[32 * g:32 * g + 32] for g in range(world)], backend.keccak)
    transcript.append(top[0])
    return ShardedCommit(backend, group, local, block, top, m, n_cols_total, shards)
