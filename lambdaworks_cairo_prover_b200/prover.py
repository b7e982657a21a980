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


def get_trace_evaluations(commits, x, frame_offsets, trace_length, ctx=None):
    """Frame::get_trace_evaluations (src/starks/frame.rs:67-83) from the trace polynomials resident in
    `commits` (the round-1 handles, in column order): row k = [t_j(x * g^offset_k) for every column j].
    Returns uint64[K, total_cols, 4]."""
    ctx = ctx or commits[0].ctx
    p = felt.MODULUS
    order = trace_length.bit_length() - 1
    g = pow(_TWO_ADIC_ROOT, 1 << (192 - order), p)
    xi = felt.to_int(x)
    points = felt.from_ints([xi * pow(g, int(k), p) % p for k in frame_offsets])
    total = sum(c.n_cols for c in commits)
    out = np.empty((len(frame_offsets), total, 4), dtype=np.uint64)
    off = 0
    for c in commits:
        ctx.check(N.lib().s252_commit_evaluate_at(c.handle, N.ptr(points), len(frame_offsets), N.ptr(out), total, off))
        off += c.n_cols
    return out


def evaluate_at(commit, point):
    """poly.evaluate(point) for every polynomial of a commit (H1(z^2), H2(z^2): prover.rs:296-300)."""
    pt = N.fe_array(np.asarray(point, dtype=np.uint64).reshape(1, 4))
    out = np.empty((1, commit.n_cols, 4), dtype=np.uint64)
    commit.ctx.check(N.lib().s252_commit_evaluate_at(commit.handle, N.ptr(pt), 1, N.ptr(out), commit.n_cols, 0))
    return out[0]


_TWO_ADIC_ROOT = 0x5282db87529cfa3f0464519c8b0fa5ad187148e11a61616070024f42f8ef94   # order 2^192


def fri_commit_phase_deep(number_layers, trace_commits, composition_commit, z, transition_offsets, trace_ood,
                          h1_z2, h2_z2, gamma, gamma_p, trace_gammas, transcript, coset_offset_u64):
    """Round 4 on the GPU (src/starks/prover.rs:327-404 after the challenges are sampled): the DEEP
    composition polynomial is built as evaluations on the LDE coset from the resident commits
    (replacing compute_deep_composition_poly, prover.rs:410-482) and handed to fri_commit_phase.
    Returns (last_value, fri_layers)."""
    from .fri import FriLayers
    ctx = trace_commits[0].ctx
    handles = (C.c_void_p * len(trace_commits))(*[c.handle for c in trace_commits])
    offs = np.ascontiguousarray(transition_offsets, dtype=np.uint64)
    ood = N.fe_array(np.asarray(trace_ood, dtype=np.uint64).reshape(-1, 4))
    gam = N.fe_array(np.asarray(trace_gammas, dtype=np.uint64).reshape(-1, 4))
    h = C.c_void_p()
    last = np.empty(4, dtype=np.uint64)
    roots = np.empty((max(number_layers, 1), 32), dtype=np.uint8)
    m = composition_commit.n_rows
    ctx.check(N.lib().s252_fri_commit_phase_deep(ctx.handle, number_layers, handles, len(trace_commits), composition_commit.handle,
                                                 N.ptr(N.fe_array(z)), N.ptr(offs), len(offs), N.ptr(ood), N.ptr(N.fe_array(h1_z2)),
                                                 N.ptr(N.fe_array(h2_z2)), N.ptr(N.fe_array(gamma)), N.ptr(N.fe_array(gamma_p)),
                                                 N.ptr(gam), transcript.handle, coset_offset_u64, C.byref(h), N.ptr(last), N.ptr(roots)))
    return last, FriLayers(ctx, h, m, roots[:number_layers])
