"""interpolate_and_commit of ONE trace on the GPUs of one box through the library's own collective entry points
(include/stark252_b200.h, "ONE trace committed by the GPUs of one box"): NCCL is called from C++, this module only passes pointers.
The torch.distributed orchestration of distributed.py does the same from Python (and runs under gloo for the CPU tests); a
non-Python caller binds the C functions used here.

    id = unique_id() on rank 0, handed to the other ranks by any means
    comm = Communicator(ctx, id, rank, world)
    sc = interpolate_and_commit_sharded(group_tables, n_rows, n_cols_total, blowup, coset_offset, comm)   # collective
    sc.root, sc.open(indices)                                                                             # collective
"""
import ctypes as C
import importlib.util
import os

import numpy as np

from . import _native as N

ID_BYTES = 128


def _prefer_bundled_nccl():
    """A process gets ONE libnccl.so.2 (the loader matches by soname).  If this interpreter has torch's bundled NCCL, bind that
    one, so that importing torch later in the same process does not find an older system NCCL already loaded under its name."""
    if os.environ.get("S252_NCCL_LIB"):
        return
    try:
        spec = importlib.util.find_spec("nvidia.nccl")
        for base in (spec.submodule_search_locations if spec and spec.submodule_search_locations else []):
            cand = os.path.join(base, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["S252_NCCL_LIB"] = cand
                return
    except Exception:
        pass


_prefer_bundled_nccl()


def unique_id():
    """ncclGetUniqueId: 128 bytes that every rank passes to Communicator."""
    buf = (C.c_uint8 * ID_BYTES)()
    if N.lib().s252_comm_unique_id(buf) != 0:
        raise N.Stark252Error(-2, "NCCL is not available (libnccl.so.2 not found; set S252_NCCL_LIB)")
    return bytes(buf)


class Communicator:
    def __init__(self, ctx, uid, rank, world):
        self.ctx = ctx
        h = C.c_void_p()
        buf = (C.c_uint8 * ID_BYTES).from_buffer_copy(uid)
        ctx.check(N.lib().s252_comm_create(ctx.handle, buf, rank, world, C.byref(h)))
        self.handle, self.rank, self.world = h, rank, world

    def close(self):
        if self.handle:
            N.lib().s252_comm_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def my_columns(n_cols_total, world, rank):
    """The contiguous column range of a rank (33 over 8 -> 5,4,4,..)."""
    base, extra = divmod(n_cols_total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class ShardedCommitHandle:
    def __init__(self, comm, handle, root):
        self.comm, self.handle, self.root = comm, handle, bytes(root)
        L = N.lib()
        self.n_rows, self.n_cols = L.s252_sharded_commit_n_rows(handle), L.s252_sharded_commit_n_cols(handle)

    def open(self, indices):
        """-> (rows uint64[n, n_cols, 4], paths uint8[n, log2(n_rows), 32] leaf -> root), on every rank."""
        idx = np.ascontiguousarray(np.asarray(indices, dtype=np.uint64))
        depth = self.n_rows.bit_length() - 1
        rows = np.zeros((len(idx), self.n_cols, 4), dtype=np.uint64)
        paths = np.zeros((len(idx), max(depth, 1), 32), dtype=np.uint8)
        self.comm.ctx.check(N.lib().s252_sharded_commit_open(self.handle, N.ptr(idx), len(idx), N.ptr(rows), N.ptr(paths)))
        return rows, paths[:, :depth]

    def free(self):
        if self.handle:
            N.lib().s252_sharded_commit_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def interpolate_and_commit_sharded(group_tables, n_rows, n_cols_total, blowup, coset_offset, comm, device_pointers=None):
    """group_tables: this rank's pipeline groups, each a C-contiguous uint64 array [n_rows, cols_g, 4] (row-major TraceTable of
    the group's columns, LW elements; pinned host memory uploads faster) -- or, with device_pointers = [(ptr, cols_g), ..], tables
    already resident on this rank's GPU."""
    L = N.lib()
    if device_pointers is not None:
        ptrs = (C.c_void_p * len(device_pointers))(*[p for p, _ in device_pointers])
        cols = (C.c_size_t * len(device_pointers))(*[c for _, c in device_pointers])
        n_groups, mem = len(device_pointers), N.DEVICE
    else:
        keep = [np.ascontiguousarray(t) if isinstance(t, np.ndarray) else t for t in group_tables]
        ptrs = (C.c_void_p * len(keep))(*[t.ctypes.data if isinstance(t, np.ndarray) else t.data_ptr() for t in keep])
        cols = (C.c_size_t * len(keep))(*[t.shape[1] for t in keep])
        n_groups, mem = len(keep), N.HOST
    h = C.c_void_p()
    root = np.zeros(32, dtype=np.uint8)
    comm.ctx.check(L.s252_interpolate_and_commit_sharded(comm.ctx.handle, comm.handle, ptrs, cols, n_groups, n_rows, n_cols_total, blowup,
                                                         coset_offset, mem, C.byref(h), N.ptr(root)), N.FFTError)
    return ShardedCommitHandle(comm, h, root.tobytes())


def generate_cairo_proof_sharded(trace, proof_options, comm, pipeline_groups=0):
    """generate_cairo_proof (src/cairo/air.rs:1183-1190) as ONE collective call over the GPUs of `comm` (s252_cairo_prove_sharded:
    the whole orchestration and NCCL inside the library).  trace: the same MainTrace on every rank.  Returns StarkProof::serialize
    bytes on rank 0, None on the other ranks."""
    out, n = C.c_void_p(), C.c_size_t()
    comm.ctx.check(N.lib().s252_cairo_prove_sharded(comm.ctx.handle, comm.handle, trace.handle, proof_options.blowup_factor,
                                                    proof_options.fri_number_of_queries, proof_options.coset_offset,
                                                    proof_options.grinding_factor, pipeline_groups, C.byref(out), C.byref(n)))
    if not out.value:
        return None
    try:
        return C.string_at(out.value, n.value)
    finally:
        N.lib().s252_cairo_proof_free(out)
