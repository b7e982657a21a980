"""generate_cairo_proof (src/cairo/air.rs:1183-1190) for ONE trace on the GPUs of one box.

One process per GPU (torch.distributed / NCCL as plumbing; the kernels are the library's own, through the
building blocks of include/stark252_cairo.h).  What is sharded and how:

  round 1   columns of the main trace are split over the ranks: every rank uploads, interpolates and
            extends only its columns, the exchange of distributed.py turns them into row blocks, every
            rank hashes its rows and builds that subtree, the roots are gathered.  The auxiliary trace is
            built REDUNDANTLY on every rank (1 ms; the 11 main columns it needs are broadcast over NVLink
            by the ranks that hold them) and its 18 columns are then sharded the same way.
  round 2   every rank evaluates the constraints on its row block (the frame's next row for the last
            `blowup` rows comes from the next rank: a 7 KB halo); the evaluations are all-gathered; H is one
            polynomial of 2N coefficients, so every rank interpolates it and extends H1/H2 itself (2 columns:
            cheaper than shipping the LDE), but the TREE over the rows of (H1, H2) is sharded like every other
            tree: each rank hashes its row block, the subtree roots are gathered (prover.rs:254-276).
  round 3   every rank evaluates ITS polynomials at z and z*g; the values are all-gathered.
  round 4   every rank builds the DEEP composition polynomial on its row block, which is its block of FRI layer 0:
            the commit phase runs sharded (fri_distributed.py: row-block trees, one pairwise exchange of half a
            layer per fold, collapse to one GPU once the layers are small), grinding is split over the ranks
            (MIN all-reduce), and every opening is served by the rank that owns the row.
Every rank runs the same Fiat-Shamir transcript, so no challenge is ever sent.
The proof bytes are identical to the single-GPU prover's (tests/test_distributed.py).
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _native as N
from . import distributed as D
from . import felt
from . import fri_distributed as F
from .merkle import DeviceCommit
from .transcript import DefaultTranscript, transcript_to_field, transcript_to_usize

P = felt.MODULUS
_TWO_ADIC_ROOT = 0x5282db87529cfa3f0464519c8b0fa5ad187148e11a61616070024f42f8ef94   # order 2^192


def _dev_tensor(ptr, n_words, device):
    return torch.as_tensor(D._DevicePointer(ptr, n_words), device=device)


def _wrap_lde(ctx, h, device):
    L = N.lib()
    commit = DeviceCommit.__new__(DeviceCommit)
    commit.ctx, commit.handle, commit.root = ctx, h, b""
    commit.n_cols, commit.n_rows, commit.n_coeffs = L.s252_commit_n_cols(h), L.s252_commit_n_rows(h), L.s252_commit_n_coeffs(h)
    ctx.adopt(commit)
    t = _dev_tensor(L.s252_commit_device_lde(h), commit.n_cols * commit.n_rows * 4, device)
    return commit, t.view(commit.n_cols, commit.n_rows, 4)


def _u64be(v):
    return int(v).to_bytes(8, "big")


def _path(p):
    return _u64be(len(p)) + b"".join(bytes(x) for x in p)


def _blob(b):
    return _u64be(len(b)) + b


class GpuCairoBackend(D.GpuBackend):
    """The compute steps of the sharded prover on this rank's GPU, through the C ABI.  Tensors hold field
    elements in the library's internal format; everything that crosses to the host (roots, challenges,
    out-of-domain values, opened rows) is in the reference's LW layout.  The tests replace this class by a
    CPU test double with the same interface, so that the orchestration below also runs under gloo."""

    def __init__(self, ctx):
        super().__init__(ctx)
        self.L = N.lib()

    # ---- transcript (host)
    def transcript(self):
        return DefaultTranscript()

    def to_field(self, t):
        return transcript_to_field(t)

    def to_usize(self, t):
        return transcript_to_usize(t)

    # ---- round 1
    def main_lde(self, trace, lo, hi, opts):
        """Upload + interpolate + extend main-trace columns [lo, hi): -> (handle, lde[c, M, 4], trace[c, N, 4])."""
        ctx, L, n = self.ctx, self.L, trace.n_rows()
        L.s252_cairo_trace_pin(trace.handle)
        cols_ptr = L.s252_cairo_trace_columns(trace.handle)
        hnd = C.c_void_p()
        ctx.check(L.s252_lde_host_columns(ctx.handle, C.c_void_p(cols_ptr + lo * n * 32), n, hi - lo, opts.blowup_factor, opts.coset_offset,
                                          1, C.byref(hnd)), N.FFTError)
        commit, lde = _wrap_lde(ctx, hnd, self.device)
        tr = _dev_tensor(L.s252_commit_device_trace(hnd), (hi - lo) * n * 4, self.device).view(hi - lo, n, 4)
        return commit, lde, tr

    def new_tensor(self, shape):
        return torch.empty(shape, dtype=torch.int64, device=self.device)

    def aux_trace(self, trace, aux_in, rap):
        """build_auxiliary_trace from trace columns 19..29 (aux_in[11, N, 4]) -> aux columns [18, N, 4]."""
        ctx, L, n = self.ctx, self.L, trace.n_rows()
        aux_ptr = C.c_void_p()
        ctx.check(L.s252_cairo_aux_trace_device(ctx.handle, trace.handle, N.ptr(rap), C.c_void_p(aux_in.data_ptr()), 1, C.byref(aux_ptr)))
        t = _dev_tensor(aux_ptr.value, 18 * n * 4, self.device).view(18, n, 4)
        t._s252_ptr = aux_ptr.value
        return t

    def free_tensor(self, t):
        if getattr(t, "_s252_ptr", None):
            self.ctx.device_free(t._s252_ptr)

    def cols_lde(self, cols, opts):
        """Interpolate + extend device columns [c, N, 4] -> (handle, lde[c, M, 4])."""
        ctx, L = self.ctx, self.L
        hnd = C.c_void_p()
        ctx.check(L.s252_lde_device_columns(ctx.handle, C.c_void_p(cols.data_ptr()), cols.shape[1], cols.shape[0], opts.blowup_factor,
                                            opts.coset_offset, C.byref(hnd)), N.FFTError)
        return _wrap_lde(ctx, hnd, self.device)

    def evaluate_at(self, handle, points):
        """poly.evaluate(point) for every polynomial of a handle: -> uint64[n_points, n_cols, 4]."""
        o = np.empty((len(points), handle.n_cols, 4), dtype=np.uint64)
        self.ctx.check(self.L.s252_commit_evaluate_at(handle.handle, N.ptr(points), len(points), N.ptr(o), handle.n_cols, 0))
        return o

    # ---- round 2
    def constraints_rows(self, trace, mblock, ablock, mhalo, ahalo, row0, rap, bco, tco, opts, out):
        rows, b = mblock.shape[1], opts.blowup_factor
        self.ctx.check(self.L.s252_cairo_constraints_rows(
            self.ctx.handle, trace.handle, C.c_void_p(mblock.data_ptr()), C.c_void_p(ablock.data_ptr()), rows, row0, rows,
            C.c_void_p(mhalo.data_ptr()), C.c_void_p(ahalo.data_ptr()), b, N.ptr(rap), N.ptr(bco), N.ptr(tco), b, opts.coset_offset,
            C.c_void_p(out.data_ptr())))

    def composition_commit(self, evals, n, opts, comp_lde_out):
        """H from its evaluations, H1/H2 LDE + tree: -> (handle, root); the LDE [2, M, 4] is copied into comp_lde_out."""
        ctx, L = self.ctx, self.L
        hnd = C.c_void_p()
        root = np.empty(32, dtype=np.uint8)
        ctx.check(L.s252_cairo_composition_commit(ctx.handle, C.c_void_p(evals.data_ptr()), n, opts.blowup_factor, opts.coset_offset,
                                                  C.byref(hnd), N.ptr(root)))
        comp = DeviceCommit(ctx, hnd, root.tobytes())
        m = n * opts.blowup_factor
        comp_lde_out.copy_(_dev_tensor(L.s252_commit_device_lde(hnd), 2 * m * 4, self.device).view(2, m, 4))
        return comp, root.tobytes()

    # ---- round 4
    def deep_rows(self, mblock, ablock, comp_lde, row0, n, z, ood, hz, gamma, gamma_p, tg, opts, out):
        rows, m = mblock.shape[1], comp_lde.shape[1]
        cblock = comp_lde[:, row0:row0 + rows]                                    # stride m
        tables = (C.c_void_p * 3)(mblock.data_ptr(), ablock.data_ptr(), cblock.data_ptr())
        strides = (C.c_size_t * 3)(rows, rows, m)
        ncs = (C.c_size_t * 3)(mblock.shape[0], ablock.shape[0], 2)
        offs = np.array([0, 1], dtype=np.uint64)
        ood_flat = np.ascontiguousarray(ood.reshape(-1, 4))
        self.ctx.check(self.L.s252_deep_rows(self.ctx.handle, tables, strides, ncs, 3, row0, rows, m, n, N.ptr(felt.from_int(z)), N.ptr(offs), 2,
                                             N.ptr(ood_flat), N.ptr(np.ascontiguousarray(hz[0])), N.ptr(np.ascontiguousarray(hz[1])),
                                             N.ptr(gamma), N.ptr(gamma_p), N.ptr(tg), opts.coset_offset, C.c_void_p(out.data_ptr())))

    def composition_lde(self, evals, n, opts):
        """H from its evaluations, H1/H2 coefficients + LDE, no tree: -> (handle, lde[2, M, 4] view of the handle's memory)."""
        hnd = C.c_void_p()
        self.ctx.check(self.L.s252_cairo_composition_lde(self.ctx.handle, C.c_void_p(evals.data_ptr()), n, opts.blowup_factor,
                                                         opts.coset_offset, C.byref(hnd)))
        return _wrap_lde(self.ctx, hnd, self.device)

    # ---- FRI building blocks (fri_distributed.py)
    def fri_fold_rows(self, v, s, i0, size, domain_size, k, zeta, coset_offset, out):
        self.ctx.check(self.L.s252_fri_fold_rows(self.ctx.handle, C.c_void_p(v.data_ptr()), C.c_void_p(s.data_ptr()), v.shape[0], i0, size,
                                                 domain_size, k, N.ptr(np.ascontiguousarray(zeta)), coset_offset, C.c_void_p(out.data_ptr())))

    def fri_continue(self, evals, n_layers, t, coset_offset, k, domain_size):
        """The rest of the commit phase on this GPU from layer k (given in full): -> (fri handle, last value LW, roots)."""
        fri = C.c_void_p()
        last = np.empty(4, dtype=np.uint64)
        roots = np.empty((max(n_layers, 1), 32), dtype=np.uint8)
        self.ctx.check(self.L.s252_fri_commit_phase_from_layer(self.ctx.handle, n_layers, C.c_void_p(evals.data_ptr()), evals.shape[0], t.handle,
                                                               coset_offset, k, C.byref(fri), N.ptr(last), N.ptr(roots)))
        return fri, last, roots[:n_layers]

    def release_fri(self, fri):
        self.L.s252_fri_destroy(fri)

    def grind_round(self, challenge, factor, base, part, parts, window_log):
        found = C.c_uint64()
        ch = np.frombuffer(challenge, dtype=np.uint8).copy()
        self.ctx.check(self.L.s252_grind_round(self.ctx.handle, N.ptr(ch), factor, base, 0, part, parts, window_log, C.byref(found)))
        return int(found.value)

    @staticmethod
    def to_bytes_be(v):
        return felt.to_bytes_be(v)

    def serialize_proof(self, n, roots, ood, hz, layers, fri_roots, last, depth, fq, opened, nonce):
        """StarkProof::serialize in the library (s252_cairo_serialize_proof) from the gathered pieces."""
        cont = lambda a, dt: np.ascontiguousarray(a, dtype=dt)                                       # noqa: E731
        q = 0 if fq is None else fq[0].shape[0]
        if q:
            ev, evs, pa, pas = (cont(fq[0], np.uint64), cont(fq[1], np.uint64), cont(fq[2], np.uint8), cont(fq[3], np.uint8))
            (mr, mp), (ar, ap), (cr, cp) = [(cont(r_, np.uint64), cont(p_, np.uint8)) for r_, p_ in opened]
            args = [N.ptr(evs), N.ptr(ev), N.ptr(pas), N.ptr(pa), N.ptr(cr), N.ptr(cp), N.ptr(mr), mr.shape[1], N.ptr(mp), N.ptr(ar), ar.shape[1], N.ptr(ap)]
            main_cols, aux_cols = mr.shape[1], ar.shape[1]
        else:
            main_cols = ood.shape[1] - 18
            args = [None, None, None, None, None, None, None, main_cols, None, None, 18, None]
        ood_c, hz_c, last_c = cont(ood, np.uint64), cont(hz, np.uint64), cont(last, np.uint64)
        fr = np.frombuffer(b"".join(fri_roots), dtype=np.uint8) if layers else np.zeros(32, dtype=np.uint8)
        rts = [np.frombuffer(r, dtype=np.uint8) for r in roots]
        out, ln = C.c_void_p(), C.c_size_t()
        rc = self.L.s252_cairo_serialize_proof(n, N.ptr(rts[0]), N.ptr(rts[1]), N.ptr(rts[2]), N.ptr(ood_c), ood.shape[1], N.ptr(hz_c), layers, N.ptr(fr),
                                               N.ptr(last_c), q, depth, *args, nonce, C.byref(out), C.byref(ln))
        if rc != N.OK:
            raise N.Stark252Error(rc, self.L.s252_cairo_last_error().decode())
        try:
            return C.string_at(out.value, ln.value)
        finally:
            self.L.s252_cairo_proof_free(out)

    def fri_commit_phase_evals(self, p0, layers, t, opts):
        """-> (fri handle, last value LW, roots uint8[layers, 32])"""
        fri = C.c_void_p()
        last = np.empty(4, dtype=np.uint64)
        roots = np.empty((max(layers, 1), 32), dtype=np.uint8)
        self.ctx.check(self.L.s252_fri_commit_phase_evals(self.ctx.handle, layers, C.c_void_p(p0.data_ptr()), p0.shape[0], t.handle,
                                                          opts.coset_offset, C.byref(fri), N.ptr(last), N.ptr(roots)))
        return fri, last, roots[:layers]

    def grind(self, challenge, factor):
        nonce = C.c_uint64()
        ch = np.frombuffer(challenge, dtype=np.uint8).copy()
        self.ctx.check(self.L.s252_generate_nonce_with_grinding(self.ctx.handle, N.ptr(ch), factor, 0, C.byref(nonce)))
        return int(nonce.value)

    def fri_query(self, fri, idx, layers, depth):
        q = len(idx)
        ev = np.empty((q, layers, 4), dtype=np.uint64)
        evs = np.empty_like(ev)
        pa = np.empty((q, layers, depth, 32), dtype=np.uint8)
        pas = np.empty_like(pa)
        ia = np.array(idx, dtype=np.uint64)
        self.ctx.check(self.L.s252_fri_query(fri, N.ptr(ia), q, N.ptr(ev), N.ptr(evs), N.ptr(pa), N.ptr(pas), depth))
        return ev, evs, pa, pas

    def commit_open(self, comp, idx, depth):
        q = len(idx)
        rows = np.empty((q, 2, 4), dtype=np.uint64)
        paths = np.empty((q, depth, 32), dtype=np.uint8)
        ia = np.array(idx, dtype=np.uint64)
        self.ctx.check(self.L.s252_commit_open(comp.handle, N.ptr(ia), q, N.ptr(rows), N.ptr(paths)))
        return rows, paths

    def release(self, fri, comp):
        self.L.s252_fri_destroy(fri)
        comp.free()


def _bcast_bytes(data, n, device, group):
    t = torch.frombuffer(bytearray(data if data is not None else bytes(n)), dtype=torch.uint8).to(device)
    dist.broadcast(t, src=0 if group is None else dist.get_global_rank(group, 0), group=group)
    return bytes(t.cpu().numpy().tobytes())


def generate_cairo_proof_sharded(trace, options, ctx, group=None, timings=None, exchange="a2a", fri_collapse_log=None,
                                 pipeline_groups=None):
    """trace: the same MainTrace on every rank (lambdaworks_cairo_prover_b200.cairo); ctx: this rank's Context (or a
    backend object with GpuCairoBackend's interface).  Returns StarkProof::serialize() bytes on rank 0 and
    None on the other ranks.  A failure on any rank (an unsatisfied AIR, say) raises on EVERY rank.
    pipeline_groups: column groups per rank of the round-1 commits (default: 2 from 2^20 rows on, else 1).  With more than
    one group the exchange of a group (point-to-point chunks) runs under the upload + LDE of the next one."""
    be = ctx if hasattr(ctx, "main_lde") else GpuCairoBackend(ctx)
    if pipeline_groups is None:
        pipeline_groups = 2 if trace.n_rows() >= (1 << 20) else 1
    with D.backend_scope(be):
        return _prove_sharded(trace, options, be, group, timings, exchange, fri_collapse_log, pipeline_groups)


def _all_ranks_ok(step, be, group):
    """Runs a rank-local step; if it raises on any rank, every rank raises (instead of the others waiting forever in
    the next collective)."""
    err = None
    try:
        out = step()
    except Exception as e:                                   # noqa: BLE001 - re-raised below on every rank
        err, out = e, None
    if dist.get_world_size(group) > 1:
        flag = torch.tensor([1 if err is not None else 0], dtype=torch.int64, device=be.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
        if int(flag.item()) and err is None:
            err = RuntimeError("a peer rank failed in a rank-local step of the sharded prover")
    if err is not None:
        raise err
    return out


def _prove_sharded(trace, options, be, group, timings, exchange, fri_collapse_log, pipeline_groups):
    import time
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    device = be.device
    gr = (lambda r: r) if group is None else (lambda r: dist.get_global_rank(group, r))
    _t = [time.perf_counter()]

    def mark(name):
        if timings is not None:
            be.sync()
            now = time.perf_counter()
            timings[name] = timings.get(name, 0.0) + (now - _t[0]) * 1e3
            _t[0] = now
    n, c_main = trace.n_rows(), trace.n_cols
    b, h = options.blowup_factor, options.coset_offset
    m = n * b
    rows_per = m // world
    if rows_per < b or m % world:
        raise ValueError("too many ranks for this trace")
    order = n.bit_length() - 1
    g = pow(_TWO_ADIC_ROOT, 1 << (192 - order), P)
    nt = 50 if c_main > 34 else 49
    t = be.transcript()

    def sharded_commit(n_cols_total, lde_of_my_columns):
        shards = D.column_shards(n_cols_total, world)
        if min(hi - lo for lo, hi in shards) == 0:
            raise ValueError("fewer columns than ranks")
        ng = max(1, min(pipeline_groups if world > 1 else 1, min(hi - lo for lo, hi in shards)))
        ranges = [D.group_ranges(hi - lo, ng) for lo, hi in shards]
        lo, hi = shards[rank]
        sub = None if timings is None else timings.setdefault("commit_detail", {})
        return D.exchange_and_commit((lde_of_my_columns(lo + a, lo + b) for a, b in ranges[rank]), ranges, shards, m, n_cols_total, t, be, group,
                                     exchange=exchange if ng == 1 else "p2p", timings=sub)

    # ---- round 1 (prover.rs:186-224)
    kept = {}

    def main_lde(lo, hi):
        handle, lde, tr = be.main_lde(trace, lo, hi, options)
        kept.setdefault("trace", []).append((lo, hi, tr))
        return handle, lde
    sc_main = sharded_commit(c_main, main_lde)
    mark("main_commit")
    rap = np.stack([be.to_field(t) for _ in range(3)])
    # build_auxiliary_trace reads main columns 19..29: each is broadcast over NVLink by the rank that holds it on its
    # device (uploading them from the host on every rank would multiply the PCIe traffic by the number of ranks)
    aux_in = be.new_tensor((11, n, 4))
    shards_main = D.column_shards(c_main, world)
    for owner, (lo, hi) in enumerate(shards_main):              # one broadcast per owner of a run of columns 19..29
        a, z_ = max(lo, 19), min(hi, 30)
        if a >= z_:
            continue
        if owner == rank:
            for tlo, thi, tr in kept["trace"]:                  # this rank's columns, one tensor per pipeline group
                x, y = max(a, tlo), min(z_, thi)
                if x < y:
                    aux_in[x - 19:y - 19].copy_(tr[x - tlo:y - tlo])
        if world > 1:
            dist.broadcast(aux_in[a - 19:z_ - 19], src=gr(owner), group=group)
    mark("aux_inputs")
    aux_cols = _all_ranks_ok(lambda: be.aux_trace(trace, aux_in, rap), be, group)
    mark("aux_build")

    def aux_lde(lo, hi):
        return be.cols_lde(aux_cols[lo:hi], options)
    sc_aux = sharded_commit(18, aux_lde)
    be.free_tensor(aux_cols)
    del aux_in, aux_cols
    mark("aux_commit")
    # ---- round 2 (prover.rs:598-640, 226-283)
    bco = np.zeros((8, 2, 4), dtype=np.uint64)
    tco = np.zeros((nt, 2, 4), dtype=np.uint64)
    for j in range(2):
        for k in range(8):
            bco[k, j] = be.to_field(t)
    for j in range(2):
        for k in range(nt):
            tco[k, j] = be.to_field(t)
    mblock, ablock = sc_main.block_tensor, sc_aux.block_tensor                  # [cols, rows_per, 4]
    mhalo = be.new_tensor((c_main, b, 4))
    ahalo = be.new_tensor((18, b, 4))
    if world == 1:
        mhalo.copy_(mblock[:, :b])
        ahalo.copy_(ablock[:, :b])
    else:
        nxt, prv = gr((rank + 1) % world), gr((rank - 1) % world)
        msend, asend = mblock[:, :b].contiguous(), ablock[:, :b].contiguous()
        for w_ in dist.batch_isend_irecv([dist.P2POp(dist.isend, msend, prv, group), dist.P2POp(dist.isend, asend, prv, group),
                                          dist.P2POp(dist.irecv, mhalo, nxt, group), dist.P2POp(dist.irecv, ahalo, nxt, group)]):
            w_.wait()
    evals = be.new_tensor((m, 4))
    mine = evals[rank * rows_per:(rank + 1) * rows_per]
    _all_ranks_ok(lambda: be.constraints_rows(trace, mblock, ablock, mhalo, ahalo, rank * rows_per, rap, bco, tco, options, mine), be, group)
    mark("constraints")
    if world > 1:
        dist.all_gather_into_tensor(evals.view(-1), mine.reshape(-1).clone(), group=group)
    # H from its evaluations and the LDE of (H1, H2) on every rank; an H above its degree bound (the trace does not
    # satisfy the AIR) is detected here, on every rank alike
    comp, comp_lde = _all_ranks_ok(lambda: be.composition_lde(evals, n, options), be, group)
    del evals
    cblock = comp_lde[:, rank * rows_per:(rank + 1) * rows_per].contiguous()
    sc_comp = F.commit_row_block(cblock, m, t, be, group)                       # appends the root (prover.rs:635)
    comp_root = sc_comp.root
    mark("composition")
    # ---- round 3 (prover.rs:650-690)
    hinv = pow(h, -1, P)
    while True:                                                                  # sample_z_ood, transcript.rs:53-70
        z = felt.to_int(be.to_field(t))
        if pow(z * hinv % P, m, P) != 1 and pow(z, n, P) != 1:
            break
    pts = felt.from_ints([z, z * g % P])

    def local_ood(local):
        return np.concatenate([be.evaluate_at(hnd, pts) for hnd in local.handles], axis=1)
    mine_ood = (local_ood(sc_main.local), local_ood(sc_aux.local))
    hz = be.evaluate_at(comp, felt.from_ints([z * z % P]))[0]                    # H1(z^2), H2(z^2): every rank holds H1, H2
    if world > 1:
        # one all-gather of fixed-size slots (the column counts per rank differ by at most one: pad to the largest)
        sm_, sa_ = D.column_shards(c_main, world), D.column_shards(18, world)
        wm, wa = max(hi - lo for lo, hi in sm_), max(hi - lo for lo, hi in sa_)
        slot = np.zeros((2, wm + wa, 4), dtype=np.uint64)
        slot[:, :mine_ood[0].shape[1]] = mine_ood[0]
        slot[:, wm:wm + mine_ood[1].shape[1]] = mine_ood[1]
        mine_t = torch.from_numpy(slot.view(np.int64)).to(device)
        all_t = torch.empty((world,) + tuple(mine_t.shape), dtype=torch.int64, device=device)
        dist.all_gather_into_tensor(all_t.view(-1), mine_t.view(-1), group=group)
        slots = all_t.cpu().numpy().view(np.uint64)                                             # [world, 2, wm + wa, 4]
        ood = np.concatenate([slots[r, :, :hi - lo] for r, (lo, hi) in enumerate(sm_)] +
                             [slots[r, :, wm:wm + hi - lo] for r, (lo, hi) in enumerate(sa_)], axis=1)   # [2, 52, 4]
    else:
        ood = np.concatenate(mine_ood, axis=1)
    ncols = ood.shape[1]
    t.append(felt.to_bytes_be(hz[0]))
    t.append(felt.to_bytes_be(hz[1]))
    for row in ood:
        for v in row:
            t.append(felt.to_bytes_be(v))
    mark("ood")
    # ---- round 4 (prover.rs:327-404)
    gamma, gamma_p = be.to_field(t), be.to_field(t)
    tg = np.stack([be.to_field(t) for _ in range(2 * ncols)])
    p0_block = be.new_tensor((rows_per, 4))
    _all_ranks_ok(lambda: be.deep_rows(mblock, ablock, comp_lde, rank * rows_per, n, z, ood, hz, gamma, gamma_p, tg, options, p0_block), be, group)
    mark("deep")
    layers = order
    q_count = options.fri_number_of_queries if layers else 0
    fri = F.fri_commit_phase_sharded(p0_block, m, layers, t, h, be, group, collapse_log=fri_collapse_log)
    mark("fri")
    nonce = F.generate_nonce_with_grinding_sharded(t.challenge(), options.grinding_factor, be, group)
    t.append(_u64be(nonce))                                                      # prover.rs:385
    idx = [be.to_usize(t) % m for _ in range(q_count)]                           # every rank samples the same indices
    mark("grinding")
    opened = D.open_many_packed([sc_main, sc_aux, sc_comp], [idx] * 3, group) if q_count else None
    fq = F.fri_query_sharded(fri, idx, layers, be, group) if q_count else None
    mark("openings")
    proof = None
    if rank == 0 and hasattr(be, "serialize_proof"):
        proof = be.serialize_proof(n, (sc_main.root, sc_aux.root, comp_root), ood, hz, layers, fri.roots, fri.last_value, m.bit_length() - 1, fq, opened,
                                   nonce)
    elif rank == 0:
        depth = m.bit_length() - 1
        if q_count:
            ev, evs, pa, pas = fq
            (main_rows, main_paths), (aux_rows, aux_paths), (crow, comp_paths) = opened
        bb = lambda a: felt.to_bytes_be_many(np.asarray(a).view(np.uint64).reshape(-1, 4))               # elements -> wire bytes
        u8 = lambda v: np.frombuffer(_u64be(v), dtype=np.uint8)
        const = lambda v: np.tile(u8(v), (q_count, 1))
        out = _u64be(n) + _u64be(2) + sc_main.root + sc_aux.root                 # StarkProof::serialize, proof/stark.rs:161-218
        frame = _u64be(2 * ncols) + _u64be(32) + bb(ood).tobytes() + _u64be(ncols)
        out += _blob(frame) + comp_root + _u64be(32) + bb(hz).tobytes()
        out += _u64be(layers) + b"".join(fri.roots) + bb(fri.last_value).tobytes() + _u64be(q_count)
        if q_count:
            # every query's blob has the same layout, so all of them are assembled as rows of one byte matrix
            def path_section(paths):                                             # [Q, layers, depth, 32] -> u64(len_k) || path_k for every layer
                offs = np.cumsum([0] + [8 + 32 * (depth - k) for k in range(layers)])
                sec = np.empty((q_count, offs[-1]), dtype=np.uint8)
                for k in range(layers):
                    sec[:, offs[k]:offs[k] + 8] = u8(depth - k)
                    sec[:, offs[k] + 8:offs[k + 1]] = paths[:, k, :depth - k].reshape(q_count, -1)
                return sec
            dec = np.concatenate([const(layers), path_section(pas), const(32), const(layers), bb(evs).reshape(q_count, -1),
                                  const(layers), bb(ev).reshape(q_count, -1), const(layers), path_section(pa)], axis=1)
            out += np.concatenate([const(dec.shape[1]), dec], axis=1).tobytes()
        out += _u64be(q_count)
        if q_count:
            as_u8 = lambda paths: np.ascontiguousarray(paths).reshape(q_count, -1)
            rows52 = np.concatenate([main_rows, aux_rows], axis=1)
            opn = np.concatenate([const(depth), as_u8(comp_paths), const(32), bb(crow).reshape(q_count, -1), const(2),
                                  const(depth), as_u8(main_paths), const(depth), as_u8(aux_paths), const(ncols),
                                  bb(rows52).reshape(q_count, -1)], axis=1)
            out += np.concatenate([const(opn.shape[1]), opn], axis=1).tobytes()
        out += _u64be(nonce)
        proof = out
    mark("serialize")
    fri.free(be)
    comp.free()
    sc_comp.free()
    sc_main.free()
    sc_aux.free()
    mark("free")
    return proof
