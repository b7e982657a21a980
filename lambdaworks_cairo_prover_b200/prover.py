"""The commitment steps of the STARK prover (src/starks/prover.rs:96-185, 254-276) on the GPU."""
import ctypes as C

import numpy as np

from . import _native as N
from . import felt
from .merkle import DeviceCommit


class TraceTable:
    """src/starks/trace.rs:9-13: row-major `table` of n_rows x n_cols elements."""

    def __init__(self, table, n_cols):
        self.table = N.fe_array(np.asarray(table, dtype=np.uint64).reshape(-1, 4))
        self.n_cols = n_cols

    def n_rows(self):
        return 0 if self.n_cols == 0 else self.table.shape[0] // self.n_cols

    @staticmethod
    def new_from_cols(cols):
        cols = np.asarray(cols, dtype=np.uint64)
        return TraceTable(np.ascontiguousarray(cols.transpose(1, 0, 2)).reshape(-1, 4), cols.shape[0])


class Domain:
    """src/starks/domain.rs:20-56 -- only the fields the commitment path reads."""

    def __init__(self, trace_length, options):
        self.blowup_factor = options.blowup_factor
        self.coset_offset = felt.from_int(options.coset_offset)
        self.coset_offset_u64 = options.coset_offset
        self.interpolation_domain_size = trace_length
        self.root_order = trace_length.bit_length() - 1
        self.lde_root_order = (trace_length * options.blowup_factor).bit_length() - 1


def interpolate_and_commit(trace, domain, transcript, ctx=None):
    """src/starks/prover.rs:126-159.  Returns the device-resident commit (trace polynomials, LDE
    columns, batched Merkle tree) and its root; the root is appended to the transcript."""
    ctx = ctx or N.default_context()
    h = C.c_void_p()
    root = np.empty(32, dtype=np.uint8)
    ctx.check(N.lib().s252_interpolate_and_commit(ctx.handle, N.ptr(trace.table), trace.n_rows(), trace.n_cols,
                                                  domain.blowup_factor, domain.coset_offset_u64, N.HOST, C.byref(h),
                                                  N.ptr(root)), N.FFTError)
    commit = DeviceCommit(ctx, h, root.tobytes())
    transcript.append(commit.root)     # prover.rs:151
    return commit, commit.root


def lde_and_commit(polys, domain, ctx=None):
    """Round 2 (src/starks/prover.rs:254-276): LDE of each polynomial + batch_commit of the zipped rows.
    polys: list of Polynomial (coefficient length <= trace length)."""
    ctx = ctx or N.default_context()
    n = max((p.coeff_len() for p in polys), default=0)
    n = max(n, 1)
    buf = np.zeros((len(polys), n, 4), dtype=np.uint64)
    for j, p in enumerate(polys):
        buf[j, :p.coeff_len()] = p.coefficients
    h = C.c_void_p()
    root = np.empty(32, dtype=np.uint8)
    ctx.check(N.lib().s252_lde_and_commit(ctx.handle, N.ptr(buf), n, len(polys), domain.interpolation_domain_size,
                                          domain.blowup_factor, domain.coset_offset_u64, N.HOST, C.byref(h), N.ptr(root)),
              N.FFTError)
    commit = DeviceCommit(ctx, h, root.tobytes())
    return commit, commit.root
