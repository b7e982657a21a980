"""Column-sharded interpolate_and_commit across the GPUs of one box.

One process per GPU.  Columns are independent for interpolation and LDE, so rank g transforms its
own contiguous range of columns with no communication.  Leaves hash whole rows, so one exchange is
needed: an all-to-all that turns "my columns, all rows" into "all columns, my row block"
(rank g ends up with rows [g*M/G, (g+1)*M/G)).  It is issued as grouped point-to-point chunks, one
per (column, destination), so that nothing is packed or staged, and per pipeline group of columns,
so that the exchange of one group overlaps the upload and LDE of the next.  Each rank hashes its block and builds that subtree;
because G is a power of two the subtree roots are exactly level log2(G) of the reference's heap,
so an all-gather of G digests and G-1 host hashes finish the tree.  The result is the same root and
the same authentication paths as the single-GPU (and the reference's) tree.

torch.distributed is the plumbing (NCCL over NVLink on GPUs; gloo in the CPU tests, where the two
compute steps are injected).  The compute steps are the library's own kernels through the C ABI.

Ordering between the library's kernels and the collectives is by stream, never by a device-wide synchronisation:
the orchestration runs with the library's stream as torch's current stream (GpuBackend.scope()), so NCCL waits for
the kernels queued before it and the kernels queued after it wait for NCCL, while the host runs ahead.
"""
import contextlib
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


def group_ranges(n_cols, n_groups):
    """Contiguous sub-ranges of one rank's columns: the units of the LDE -> exchange pipeline."""
    n_groups = max(1, min(n_groups, n_cols))
    return column_shards(n_cols, n_groups)


class _LocalColumns:
    """This rank's columns (coefficients + LDE, all rows), held as one handle per pipeline group."""

    def __init__(self, handles, ranges):
        self.handles, self.ranges = handles, ranges

    def _find(self, j):
        for h, (lo, hi) in zip(self.handles, self.ranges):
            if lo <= j < hi:
                return h, j - lo
        raise IndexError(j)

    def coefficients(self, j):
        h, k = self._find(j)
        return h.coefficients(k)

    def lde_column(self, j, first=0, count=None):
        h, k = self._find(j)
        return h.lde_column(k, first, count)

    def free(self):
        for h in self.handles:
            if h is not None and hasattr(h, "free"):
                h.free()


class GpuBackend:
    """The two compute steps on this rank's GPU, through the C ABI."""

    def __init__(self, ctx):
        self.ctx = ctx
        self.device = torch.device("cuda", ctx.device)
        self.stream = torch.cuda.ExternalStream(ctx.stream, device=self.device)
        self._staging = {}

    def scope(self):
        """Everything torch does inside (copies, allocations, NCCL collectives) is ordered on the library's stream."""
        return torch.cuda.stream(self.stream)

    def _wrap(self, h, n_rows, n_cols, blowup):
        commit = DeviceCommit.__new__(DeviceCommit)
        commit.ctx, commit.handle, commit.root = self.ctx, h, b""
        commit.n_cols, commit.n_rows, commit.n_coeffs = n_cols, n_rows * blowup, n_rows
        self.ctx.adopt(commit)
        ptr = N.lib().s252_commit_device_lde(h)
        t = torch.as_tensor(_DevicePointer(ptr, n_cols * n_rows * blowup * 4), device=self.device)
        return commit, t.view(n_cols, n_rows * blowup, 4)

    def lde(self, shard_table, n_rows, n_cols, blowup, coset_offset):
        """-> (handle, tensor[n_cols, M, 4] int64 view of the LDE columns)"""
        h = C.c_void_p()
        self.ctx.check(N.lib().s252_interpolate_and_lde(self.ctx.handle, N.ptr(shard_table), n_rows, n_cols, blowup,
                                                        coset_offset, N.HOST, C.byref(h)), N.FFTError)
        return self._wrap(h, n_rows, n_cols, blowup)

    def lde_pipeline(self, group_tables, n_rows, blowup, coset_offset):
        """Yields (handle, lde tensor) per group.  The upload of group g+1 runs on the library's copy stream
        while group g is interpolated and extended (the tables should be pinned host memory)."""
        if all(getattr(t, "is_cuda", False) for t in group_tables):
            # the groups are already resident in this GPU's HBM: no staging, no upload
            for t in group_tables:
                h = C.c_void_p()
                self.ctx.check(N.lib().s252_interpolate_and_lde(self.ctx.handle, C.c_void_p(t.data_ptr()), n_rows, t.shape[-2], blowup,
                                                                coset_offset, N.DEVICE, C.byref(h)), N.FFTError)
                yield self._wrap(h, n_rows, t.shape[-2], blowup)
            return
        sizes = [t.nbytes for t in group_tables]
        key = tuple(sizes)
        if key not in self._staging:                      # device staging, reused by later commits of the same shape
            for bufs in self._staging.values():
                for p in bufs:
                    self.ctx.device_free(p)
            self._staging = {key: [self.ctx.device_alloc(sz) for sz in sizes]}
        staging = self._staging[key]

        def upload(g):
            t = group_tables[g]
            addr = t.data_ptr() if hasattr(t, "data_ptr") else t.ctypes.data
            self.ctx.to_device_async(staging[g], addr, sizes[g])

        upload(0)
        for g, t in enumerate(group_tables):
            self.ctx.copy_stream_wait()
            if g + 1 < len(group_tables):
                upload(g + 1)
            c = t.shape[-2] if len(t.shape) == 3 else t.shape[0] // n_rows
            h = C.c_void_p()
            self.ctx.check(N.lib().s252_interpolate_and_lde(self.ctx.handle, C.c_void_p(staging[g]), n_rows, c, blowup,
                                                            coset_offset, N.DEVICE, C.byref(h)), N.FFTError)
            yield self._wrap(h, n_rows, c, blowup)

    def commit_block(self, cols):
        """cols: tensor[c_total, rows, 4] on this device -> (DeviceCommit with the subtree, subtree root bytes).
        The tree is built over `cols` in place; the returned handle keeps the tensor alive."""
        c_total, rows = cols.shape[0], cols.shape[1]
        h = C.c_void_p()
        root = np.empty(32, dtype=np.uint8)
        self.ctx.check(N.lib().s252_commit_device_columns_inplace(self.ctx.handle, C.c_void_p(cols.data_ptr()), rows, c_total, rows,
                                                                  C.byref(h), N.ptr(root)))
        commit = DeviceCommit(self.ctx, h, root.tobytes())
        commit._keepalive = cols
        return commit, root.tobytes()

    def commit_block_nosync(self, cols):
        """commit_block without the read-back: -> (DeviceCommit, int64[4] device tensor viewing the subtree root)."""
        c_total, rows = cols.shape[0], cols.shape[1]
        h = C.c_void_p()
        self.ctx.check(N.lib().s252_commit_device_columns_inplace(self.ctx.handle, C.c_void_p(cols.data_ptr()), rows, c_total, rows,
                                                                  C.byref(h), None))
        commit = DeviceCommit(self.ctx, h, b"")
        commit._keepalive = cols
        root = torch.as_tensor(_DevicePointer(N.lib().s252_commit_device_nodes(h), 4), device=self.device)
        return commit, root

    def before_collective(self):
        pass                            # stream-ordered: see scope()

    def after_collective(self):
        pass

    def sync(self):
        """Host waits for this rank's queue (timing marks only)."""
        self.ctx.synchronize()
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


def open_many(commits, indices, index_lists=None):
    """ShardedCommit.open for several commits with ONE exchange: returns [(rows, paths), ...] in the order of
    `commits`.  indices: the same positions for every commit (tables over one row partition), or index_lists: one
    list of positions per commit (FRI layers: iota mod the layer size and its symmetric)."""
    first = commits[0]
    world = dist.get_world_size(first.group)
    rank = dist.get_rank(first.group)
    if index_lists is None:
        index_lists = [indices] * len(commits)
    part = {}
    for ci, sc in enumerate(commits):
        rows_per = sc.n_rows // world
        mine = [(q, i % rows_per) for q, i in enumerate(index_lists[ci]) if i // rows_per == rank]
        if mine:
            rows, paths = sc.backend.open_block(sc.block, [i for _, i in mine])
            for k, (q, _) in enumerate(mine):
                part[(ci, q)] = (np.asarray(rows[k]).copy(), [bytes(np.asarray(p).tobytes()) for p in paths[k]])
    gathered = [None] * world
    if world > 1:
        dist.all_gather_object(gathered, part, group=first.group)
    else:
        gathered = [part]
    merged = {}
    for p in gathered:
        merged.update(p)
    out = []
    for ci, sc in enumerate(commits):
        out_rows, out_paths = [], []
        rows_per = sc.n_rows // world
        for q, i in enumerate(index_lists[ci]):
            rows, path = merged[(ci, q)]
            node = (world - 1) + i // rows_per
            while node != 0:
                sib = node + 1 if node & 1 else node - 1
                path = path + [sc.top[sib]]
                node = (node - 1) >> 1
            out_rows.append(rows)
            out_paths.append(path)
        out.append((out_rows, out_paths))
    return out


def open_many_packed(commits, index_lists, group=None):
    """The openings of several row-block commits gathered on rank 0 with ONE reduction and no per-element Python work:
    every rank writes the rows and subtree paths of the positions it owns into a zero-filled byte buffer with a fixed
    layout (an entry has exactly one owner), the buffers are summed onto rank 0, and rank 0 appends the replicated top
    levels.  Returns, on rank 0, [(rows uint64[n, n_cols, 4], paths uint8[n, depth, 32]), ...] in the order of `commits`
    (paths leaf -> root); None on the other ranks."""
    first = commits[0]
    be = first.backend
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    metas, total = [], 0
    for sc, idxs in zip(commits, index_lists):
        rows_per = sc.n_rows // world
        dl = rows_per.bit_length() - 1
        esz = 32 * (sc.n_cols + dl)
        metas.append((total, len(idxs), esz, dl, rows_per))
        total += len(idxs) * esz
    host = np.zeros(total, dtype=np.uint8)
    for sc, idxs, (off, n, esz, dl, rows_per) in zip(commits, index_lists, metas):
        idx = np.asarray(idxs, dtype=np.int64)
        mine = np.nonzero(idx // rows_per == rank)[0]
        if mine.size == 0:
            continue
        rows, paths = be.open_block(sc.block, [int(i) for i in idx[mine] % rows_per])
        view = host[off:off + n * esz].reshape(n, esz)
        view[mine, :32 * sc.n_cols] = np.stack([np.ascontiguousarray(r).view(np.uint8).reshape(-1) for r in rows])
        if dl:
            view[mine, 32 * sc.n_cols:] = np.stack([np.frombuffer(b"".join(bytes(np.asarray(x).tobytes()) for x in p_), dtype=np.uint8)
                                                    if not isinstance(p_, np.ndarray) else p_.reshape(-1) for p_ in paths])
    if world > 1:
        t = torch.from_numpy(host.view(np.int64)).to(be.device)
        dist.reduce(t, dst=0 if group is None else dist.get_global_rank(group, 0), op=dist.ReduceOp.SUM, group=group)
        if rank != 0:
            return None
        host = t.cpu().numpy().view(np.uint8)
    dt = world.bit_length() - 1
    out = []
    for sc, idxs, (off, n, esz, dl, rows_per) in zip(commits, index_lists, metas):
        view = host[off:off + n * esz].reshape(n, esz)
        rows = np.ascontiguousarray(view[:, :32 * sc.n_cols]).view(np.uint64).reshape(n, sc.n_cols, 4)
        paths = np.zeros((n, dl + dt, 32), dtype=np.uint8)
        if dl:
            paths[:, :dl] = view[:, 32 * sc.n_cols:].reshape(n, dl, 32)
        if dt:
            # the top of the tree is replicated: the siblings above rank g's subtree root (heap node world-1+g)
            top = np.zeros((world, dt, 32), dtype=np.uint8)
            for g in range(world):
                node, lvl = (world - 1) + g, 0
                while node != 0:
                    sib = node + 1 if node & 1 else node - 1
                    top[g, lvl] = np.frombuffer(sc.top[sib], dtype=np.uint8)
                    node = (node - 1) >> 1
                    lvl += 1
            paths[:, dl:] = top[np.asarray(idxs, dtype=np.int64) // rows_per]
        out.append((rows, paths))
    return out


def interpolate_and_commit_sharded(shard_table, n_rows, n_cols_total, blowup, coset_offset, transcript, backend, group=None,
                                   pipeline_groups=1, exchange="p2p"):
    """interpolate_and_commit (src/starks/prover.rs:126-159) for ONE trace whose columns are spread
    over the ranks of `group`.  shard_table: this rank's columns as a row-major table
    (n_rows x c_rank, i.e. TraceTable::get_cols, src/starks/trace.rs:31-43) -- or a LIST of such
    tables, one per pipeline group (group_ranges(c_rank, K)), in which case the upload and LDE of
    group g+1 overlap the exchange of group g.  Every rank gets the root (appended to its transcript,
    prover.rs:151).

    Exchange: rank r sends rows [d*M/G, (d+1)*M/G) of each of its columns to rank d -- one contiguous
    chunk per (column, destination), straight out of the LDE buffer and straight into the receiver's
    column-major row block (no packing, no staging copy), as grouped point-to-point operations."""
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
    tables = list(shard_table) if isinstance(shard_table, (list, tuple)) else None
    n_groups = len(tables) if tables is not None else 1
    # every rank runs the same number of exchange rounds; round k carries group k of every rank's columns
    ranges = [group_ranges(c, n_groups) for c in counts]
    if any(len(r) != n_groups for r in ranges):
        raise ValueError("more pipeline groups than columns on some rank")
    if tables is not None and hasattr(backend, "lde_pipeline"):
        producer = backend.lde_pipeline(tables, n_rows, blowup, coset_offset)
    elif tables is not None:
        producer = (backend.lde(t, n_rows, ranges[rank][g][1] - ranges[rank][g][0], blowup, coset_offset) for g, t in enumerate(tables))
    else:
        producer = iter([backend.lde(shard_table, n_rows, c_mine, blowup, coset_offset)])

    with backend_scope(backend):
        return exchange_and_commit(producer, ranges, shards, m, n_cols_total, transcript, backend, group, exchange=exchange)


def backend_scope(backend):
    return backend.scope() if hasattr(backend, "scope") else contextlib.nullcontext()


def gather_subtree_roots(backend, block_cols, group):
    """Hash this rank's row block, build its subtree and all-gather the G subtree roots: -> (block handle, 32*G bytes).  On GPUs
    the roots go device to device and the host waits once, for the gathered digests."""
    world = dist.get_world_size(group)
    if hasattr(backend, "commit_block_nosync"):
        block, mine = backend.commit_block_nosync(block_cols)
        gathered = torch.empty(4 * world, dtype=torch.int64, device=mine.device)
        if world > 1:
            dist.all_gather_into_tensor(gathered, mine, group=group)
        else:
            gathered.copy_(mine)
        roots = gathered.cpu().numpy().tobytes()
        rank = dist.get_rank(group)
        block.root = roots[32 * rank:32 * rank + 32]
        return block, roots
    block, sub_root = backend.commit_block(block_cols)
    mine = torch.frombuffer(bytearray(sub_root), dtype=torch.uint8).to(block_cols.device)
    gathered = torch.empty(32 * world, dtype=torch.uint8, device=block_cols.device)
    if world > 1:
        dist.all_gather_into_tensor(gathered, mine, group=group)
    else:
        gathered.copy_(mine)
    return block, bytes(gathered.cpu().numpy().tobytes())


def exchange_and_commit(producer, ranges, shards, m, n_cols_total, transcript, backend, group=None, exchange="p2p", timings=None):
    """The exchange + per-rank subtree + top of the tree for LDE columns produced group by group:
    `producer` yields (handle, lde[c_group, M, 4]) for this rank's pipeline groups in order; ranges[r][g] is
    the column range (within rank r's shard) of rank r's group g.
    exchange: "p2p" = one point-to-point chunk per (column, destination), nothing packed (large tables,
    pipelined groups); "a2a" = one packed all_to_all_single (one NCCL call instead of dozens: better when
    the table is small and the launch latency of the chunks would dominate; single group only)."""
    import time
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    rows_per = m // world
    recv = None
    handles, works = [], []
    clock = [time.perf_counter()]

    def mark(name):
        if timings is not None:
            if hasattr(backend, "sync"):
                backend.sync()
            now = time.perf_counter()
            timings[name] = timings.get(name, 0.0) + (now - clock[0]) * 1e3
            clock[0] = now
    for g, (handle, lde) in enumerate(producer):                 # lde: [c_group, M, 4], complete when yielded
        mark("lde")
        handles.append(handle)
        if recv is None:
            recv = torch.empty((n_cols_total, rows_per, 4), dtype=lde.dtype, device=lde.device)
        lo, hi = ranges[rank][g]
        if exchange == "a2a":
            if len(ranges[rank]) != 1:
                raise ValueError("the packed all-to-all exchange takes a single pipeline group")
            c_mine = hi - lo
            counts = [b - a for a, b in shards]
            send = lde.view(c_mine, world, rows_per * 4).permute(1, 0, 2).contiguous().view(world * c_mine, rows_per * 4)
            backend.before_collective()
            dist.all_to_all_single(recv.view(n_cols_total, rows_per * 4), send, output_split_sizes=counts,
                                   input_split_sizes=[c_mine] * world, group=group)
            continue
        ops = []
        for d in range(world):
            if d == rank:
                continue
            for c in range(hi - lo):
                ops.append(dist.P2POp(dist.isend, lde[c, d * rows_per:(d + 1) * rows_per], d if group is None else dist.get_global_rank(group, d), group))
        for src in range(world):
            if src == rank:
                continue
            slo, shi = ranges[src][g]
            for c in range(slo, shi):
                ops.append(dist.P2POp(dist.irecv, recv[shards[src][0] + c], src if group is None else dist.get_global_rank(group, src), group))
        backend.before_collective()
        recv[shards[rank][0] + lo:shards[rank][0] + hi].copy_(lde[:, rank * rows_per:(rank + 1) * rows_per])   # my own rows
        if ops:
            works.extend(dist.batch_isend_irecv(ops))
    for w in works:
        w.wait()
    backend.after_collective()
    mark("exchange")
    block, roots = gather_subtree_roots(backend, recv, group)
    mark("hash")
    top = build_top([roots[32 * g:32 * g + 32] for g in range(world)], backend.keccak)
    transcript.append(top[0])
    mark("roots")
    sc = ShardedCommit(backend, group, _LocalColumns(handles, ranges[rank]), block, top, m, n_cols_total, shards)
    sc.block_tensor = recv          # [n_cols_total, rows_per, 4]: all columns, this rank's rows
    return sc
