"""A restatement of the reference's generic STARK prover and verifier (src/starks/prover.rs,
src/starks/verifier.rs) for its toy AIRs, with the LDE + commitment hot path behind a back-end:

    OracleBackend  -> the CPU oracle (oracle/)
    GpuBackend     -> the product's C ABI (CUDA path)

Everything outside the hot path (constraint evaluation, composition polynomial, OOD evaluations, DEEP
polynomial) is plain python-integer arithmetic shared by both, so `prove` run twice must yield the
same serialized StarkProof bytes, and `verify` (pinned on the reference's golden proof for steps
1, 3, 4) must accept them.  TEST INFRASTRUCTURE: sizes are tiny (N <= 64).
"""
import numpy as np

from oracle import pyoracle as O
from oracle.proof_format import DeepPolynomialOpenings, Frame, FriDecommitment, StarkProof

P = O.P


# ------------------------------------------------------------------------------ polynomials
def poly_eval(c, x):
    acc = 0
    for v in reversed(c):
        acc = (acc * x + v) % P
    return acc


def poly_trim(c):
    c = list(c)
    while c and c[-1] == 0:
        c.pop()
    return c


def poly_add(a, b):
    n = max(len(a), len(b))
    return poly_trim([((a[i] if i < len(a) else 0) + (b[i] if i < len(b) else 0)) % P for i in range(n)])


def poly_scale(a, s):
    return poly_trim([v * s % P for v in a])


def poly_sub_const(a, v):
    a = list(a) if a else [0]
    a[0] = (a[0] - v) % P
    return poly_trim(a)


def ruffini(a, b):
    """(a(X) - a(b)) / (X - b): Polynomial::ruffini_division_inplace."""
    if not a:
        return []
    out = [0] * (len(a) - 1)
    c = 0
    for i in range(len(a) - 1, 0, -1):
        c = (a[i] + c * b) % P
        out[i - 1] = c
    return poly_trim(out)


def poly_mul_linear(a, root):
    """a(X) * (X - root)"""
    out = [0] * (len(a) + 1)
    for i, v in enumerate(a):
        out[i + 1] = (out[i + 1] + v) % P
        out[i] = (out[i] - v * root) % P
    return out


# ------------------------------------------------------------------------------ AIRs
class AIR:
    """AirContext + the callbacks of `trait AIR` (src/starks/traits.rs:15-119)."""
    trace_columns = 1
    transition_degrees = [1]
    transition_exemptions = [2]
    transition_offsets = [0, 1, 2]
    num_transition_constraints = 1
    num_transition_exemptions = 1

    def __init__(self, trace_length, pub_inputs, options):
        self.trace_length, self.pub_inputs, self.options = trace_length, pub_inputs, options

    def composition_poly_degree_bound(self):
        return self.trace_length


class FibonacciAIR(AIR):                      # src/starks/example/simple_fibonacci.rs
    def compute_transition(self, frame):
        return [(frame[2][0] - frame[1][0] - frame[0][0]) % P]

    def boundary_constraints(self):           # (col, step, value)
        return [(0, 0, self.pub_inputs[0]), (0, 1, self.pub_inputs[1])]

    @staticmethod
    def trace(a0, a1, n):
        t = [a0, a1]
        while len(t) < n:
            t.append((t[-1] + t[-2]) % P)
        return [t]


class Fibonacci2ColsAIR(AIR):                 # src/starks/example/fibonacci_2_columns.rs
    trace_columns = 2
    transition_degrees = [1, 1]
    transition_exemptions = [1, 1]
    transition_offsets = [0, 1]
    num_transition_constraints = 2

    def compute_transition(self, frame):
        f, s = frame[0], frame[1]
        return [(s[0] - f[0] - f[1]) % P, (s[1] - f[1] - s[0]) % P]

    def boundary_constraints(self):
        return [(0, 0, self.pub_inputs[0]), (1, 0, self.pub_inputs[1])]

    @staticmethod
    def trace(a0, a1, n):
        c1, c2 = [a0], [a1]
        for i in range(1, n):
            nv = (c1[i - 1] + c2[i - 1]) % P
            c1.append(nv)
            c2.append((nv + c2[i - 1]) % P)
        return [c1, c2]


class QuadraticAIR(AIR):                      # src/starks/example/quadratic_air.rs
    transition_degrees = [2]
    transition_exemptions = [1]
    transition_offsets = [0, 1]

    def composition_poly_degree_bound(self):
        return 2 * self.trace_length

    def compute_transition(self, frame):
        return [(frame[1][0] - frame[0][0] * frame[0][0]) % P]

    def boundary_constraints(self):
        return [(0, 0, self.pub_inputs[0])]

    @staticmethod
    def trace(a0, n):
        t = [a0]
        while len(t) < n:
            t.append(t[-1] * t[-1] % P)
        return [t]


# ------------------------------------------------------------------------------ back-ends
class OracleBackend:
    name = "oracle"

    def transcript(self):
        return O.Transcript()

    @staticmethod
    def to_field(t):
        return O.lw_to_int(t.to_field())

    @staticmethod
    def to_usize(t):
        return t.to_usize()

    def interpolate_and_commit(self, cols, blowup, offset):
        n, c = len(cols[0]), len(cols)
        table = O.ints_to_lw([cols[j][i] for i in range(n) for j in range(c)]).reshape(n, c, 4)
        r = O.interpolate_and_commit(table, blowup, offset)
        return _Commit([O.lw_to_ints(x) for x in r["coeffs"]], [O.lw_to_ints(x) for x in r["lde"]], r["root"],
                       lambda i: [bytes(p) for p in O.merkle_path(r["nodes"], i)])

    def lde_and_commit(self, polys, n, blowup, offset):
        ldes = [O.evaluate_polynomial_on_lde_domain(O.ints_to_lw(p) if p else np.zeros((0, 4), np.uint64), blowup, n,
                                                    O.fe_from_u64(offset)) for p in polys]
        nodes, root = O.commit_columns(np.stack(ldes))
        return _Commit(None, [O.lw_to_ints(x) for x in ldes], root, lambda i: [bytes(p) for p in O.merkle_path(nodes, i)])

    def fri_commit_phase(self, layers, p0, t, offset, m):
        last, roots, evals, nodes = O.fri_commit_phase(layers, O.ints_to_lw(p0) if p0 else np.zeros((0, 4), np.uint64), t,
                                                       O.fe_from_u64(offset), m)

        def query(iota):
            ev, ev_sym, pa, pa_sym = [], [], [], []
            for k in range(layers):
                size = m >> k
                i, isym = iota % size, (iota + size // 2) % size
                ev.append(O.lw_to_int(evals[k][i]))
                ev_sym.append(O.lw_to_int(evals[k][isym]))
                pa.append([bytes(p) for p in O.merkle_path(nodes[k], i)])
                pa_sym.append([bytes(p) for p in O.merkle_path(nodes[k], isym)])
            return FriDecommitment(pa_sym, ev_sym, ev, pa)
        return O.lw_to_int(last), [r.tobytes() for r in roots], query

    def grinding(self, challenge, factor):
        return O.generate_nonce_with_grinding(challenge, factor)


class GpuBackend:
    name = "gpu"

    def __init__(self, ctx, deep_on_device=False):
        import lambdaworks_cairo_prover_b200 as P_
        from lambdaworks_cairo_prover_b200 import felt
        self.P, self.felt, self.ctx = P_, felt, ctx
        self.deep_on_device = deep_on_device

    def ood_evaluations(self, main, comp, z, offsets, n):
        zf = self.felt.from_int(z)
        ood = self.P.get_trace_evaluations([main.keep], zf, offsets, n, self.ctx)
        hz = self.P.evaluate_at(comp.keep, self.felt.from_int(z * z % P))
        return [self.felt.to_ints(row) for row in ood], self.felt.to_int(hz[0]), self.felt.to_int(hz[1])

    def fri_deep(self, layers, main, comp, z, offsets, ood, h1z, h2z, gamma, gamma_p, tg, t, offset):
        f = self.felt
        ood_arr = np.stack([f.from_ints(row) for row in ood])
        last, fl = self.P.fri_commit_phase_deep(layers, [main.keep], comp.keep, f.from_int(z), offsets, ood_arr, f.from_int(h1z),
                                                f.from_int(h2z), f.from_int(gamma), f.from_int(gamma_p), f.from_ints(tg), t, offset)
        return f.to_int(last), [layer.root for layer in fl], self._query_fn(fl)

    def _query_fn(self, fl):
        def query(iota):
            q = self.P.fri_open(fl, [iota])[0]
            return FriDecommitment([p.merkle_path for p in q.layers_auth_paths_sym], self.felt.to_ints(np.stack(q.layers_evaluations_sym)),
                                   self.felt.to_ints(np.stack(q.layers_evaluations)), [p.merkle_path for p in q.layers_auth_paths])
        return query

    def transcript(self):
        return self.P.DefaultTranscript()

    def to_field(self, t):
        return self.felt.to_int(self.P.transcript_to_field(t))

    def to_usize(self, t):
        return self.P.transcript_to_usize(t)

    def _wrap(self, commit, with_coeffs):
        c = commit.n_cols
        coeffs = [self.felt.to_ints(commit.coefficients(j)) for j in range(c)] if with_coeffs else None
        lde = [self.felt.to_ints(commit.lde_column(j)) for j in range(c)]
        return _Commit(coeffs, lde, commit.root, lambda i: commit.get_proof_by_pos(i).merkle_path, keep=commit)

    def interpolate_and_commit(self, cols, blowup, offset):
        n, c = len(cols[0]), len(cols)
        table = self.felt.from_ints([cols[j][i] for i in range(n) for j in range(c)])
        opts = self.P.ProofOptions(blowup, 1, offset, 1)
        scratch = self.P.DefaultTranscript()      # the caller appends the root itself
        commit, _ = self.P.interpolate_and_commit(self.P.TraceTable(table, c), self.P.Domain(n, opts), scratch, self.ctx)
        return self._wrap(commit, True)

    def lde_and_commit(self, polys, n, blowup, offset):
        ps = [self.P.Polynomial(self.felt.from_ints(p) if p else np.zeros((0, 4), np.uint64)) for p in polys]
        commit, _ = self.P.lde_and_commit(ps, self.P.Domain(n, self.P.ProofOptions(blowup, 1, offset, 1)), self.ctx)
        return self._wrap(commit, False)

    def fri_commit_phase(self, layers, p0, t, offset, m):
        poly = self.P.Polynomial(self.felt.from_ints(p0) if p0 else np.zeros((0, 4), np.uint64))
        last, fl = self.P.fri_commit_phase(layers, poly, t, self.felt.from_int(offset), m, self.ctx)

        def query(iota):
            q = self.P.fri_open(fl, [iota])[0]
            return FriDecommitment([p.merkle_path for p in q.layers_auth_paths_sym], self.felt.to_ints(np.stack(q.layers_evaluations_sym)),
                                   self.felt.to_ints(np.stack(q.layers_evaluations)), [p.merkle_path for p in q.layers_auth_paths])
        return self.felt.to_int(last), [layer.root for layer in fl], query

    def grinding(self, challenge, factor):
        return self.P.generate_nonce_with_grinding(challenge, factor, self.ctx)


class _Commit:
    def __init__(self, coeffs, lde, root, path_fn, keep=None):
        self.coeffs, self.lde, self.root, self.path, self.keep = coeffs, lde, bytes(root), path_fn, keep


# ------------------------------------------------------------------------------ prover
def _felt_bytes(v):
    return int(v).to_bytes(32, "big")


def _domain(n, blowup, offset):
    m = n * blowup
    w = O.lw_to_int(O.primitive_root(m.bit_length() - 1))
    g = pow(w, blowup, P)
    lde = [offset * pow(w, i, P) % P for i in range(m)]
    roots = [pow(g, i, P) for i in range(n)]
    return w, g, lde, roots


def _exemption_polys(air, roots):
    """AIR::transition_exemptions (traits.rs:42-76): one polynomial per distinct non-zero exemption."""
    out, seen = [], []
    for e in air.transition_exemptions:
        if e > 0 and e not in seen:
            seen.append(e)
            poly = [1]
            for r in list(reversed(roots))[:e]:
                poly = poly_mul_linear(poly, r)
            out.append(poly)
    return out, seen


def prove(air_cls, trace_cols, pub_inputs, options, be):
    """src/starks/prover.rs:532-776 with the hot path delegated to `be`."""
    n = len(trace_cols[0])
    blowup, offset = options.blowup_factor, options.coset_offset
    air = air_cls(n, pub_inputs, options)
    m = n * blowup
    w, g, lde_dom, roots = _domain(n, blowup, offset)
    t = be.transcript()
    # ---- round 1
    main = be.interpolate_and_commit(trace_cols, blowup, offset)
    t.append(main.root)
    polys, lde = main.coeffs, main.lde
    ncols = len(polys)
    # ---- round 2
    bcs = air.boundary_constraints()
    b_alpha = [be.to_field(t) for _ in bcs]
    b_beta = [be.to_field(t) for _ in bcs]
    t_alpha = [be.to_field(t) for _ in range(air.num_transition_constraints)]
    t_beta = [be.to_field(t) for _ in range(air.num_transition_constraints)]
    bound = air.composition_poly_degree_bound()
    b_adj = bound - n
    ex_polys, ex_keys = _exemption_polys(air, roots)
    max_deg = max(air.transition_degrees)
    offset_pow = pow(offset, n, P)
    wb = O.lw_to_int(O.primitive_root(blowup.bit_length() - 1))
    zerofier_inv = [pow((offset_pow * pow(wb, i, P) - 1) % P, -1, P) for i in range(blowup)]
    evals = []
    for i, d in enumerate(lde_dom):
        acc = 0
        for (col, step, value), a, b in zip(bcs, b_alpha, b_beta):
            zinv = pow((d - pow(g, step, P)) % P, -1, P)
            acc += zinv * ((a * pow(d, b_adj, P) + b) % P) * ((lde[col][i] - value) % P)
        frame = [[lde[j][(i + k * blowup) % m] for j in range(ncols)] for k in air.transition_offsets]
        tr = air.compute_transition(frame)
        for ev, ex, deg, a, b in zip(tr, air.transition_exemptions, air.transition_degrees, t_alpha, t_beta):
            term = zerofier_inv[i % blowup] * ((a * pow(d, bound - n * (deg - 1), P) + b) % P) * ev
            if ex != 0:
                idx = 0 if air.num_transition_exemptions == 1 else ex_keys.index(ex)
                term *= poly_eval(ex_polys[idx], d)
            acc += term
        evals.append(acc % P)
    h = poly_trim(O.lw_to_ints(O.interpolate_offset_fft(O.ints_to_lw(evals), O.fe_from_u64(offset))))
    h1, h2 = poly_trim(h[0::2]), poly_trim(h[1::2])
    comp = be.lde_and_commit([h1, h2], n, blowup, offset)
    t.append(comp.root)
    # ---- round 3
    while True:
        z = be.to_field(t)
        if z not in lde_dom and z not in roots:
            break
    z2 = z * z % P
    if getattr(be, "deep_on_device", False):
        ood, h1z, h2z = be.ood_evaluations(main, comp, z, air.transition_offsets, n)      # round 3 on the GPU
    else:
        h1z, h2z = poly_eval(h1, z2), poly_eval(h2, z2)
        ood = [[poly_eval(p, z * pow(g, k, P) % P) for p in polys] for k in air.transition_offsets]
    t.append(_felt_bytes(h1z))
    t.append(_felt_bytes(h2z))
    for row in ood:
        for v in row:
            t.append(_felt_bytes(v))
    # ---- round 4
    gamma, gamma_p = be.to_field(t), be.to_field(t)
    tg = [be.to_field(t) for _ in range(len(air.transition_offsets) * ncols)]
    layers = n.bit_length() - 1
    if getattr(be, "deep_on_device", False):
        # round 4 on the GPU: DEEP polynomial in the evaluation domain straight into FRI
        last, fri_roots, fri_query = be.fri_deep(layers, main, comp, z, air.transition_offsets, ood, h1z, h2z, gamma, gamma_p, tg, t, offset)
    else:
        deep = poly_add(poly_scale(ruffini(poly_sub_const(h1, h1z), z2), gamma), poly_scale(ruffini(poly_sub_const(h2, h2z), z2), gamma_p))
        flen = len(air.transition_offsets)
        for i, tj in enumerate(polys):
            for r, k in enumerate(air.transition_offsets):
                zs = z * pow(g, k, P) % P
                deep = poly_add(deep, poly_scale(ruffini(poly_sub_const(tj, ood[r][i]), zs), tg[i * flen + r]))
        last, fri_roots, fri_query = be.fri_commit_phase(layers, deep, t, offset, m)
    nonce = be.grinding(t.challenge(), options.grinding_factor)
    assert nonce is not None, "nonce not found"
    t.append(nonce.to_bytes(8, "big"))
    query_list, openings = [], []
    if layers:
        iotas = [be.to_usize(t) % m for _ in range(options.fri_number_of_queries)]
        query_list = [fri_query(i) for i in iotas]
        for iota in iotas:
            idx = iota % m
            openings.append(DeepPolynomialOpenings(comp.path(idx), comp.lde[0][idx], comp.lde[1][idx], [main.path(idx)],
                                                   [lde[j][idx] for j in range(ncols)]))
    return StarkProof(n, [main.root], Frame([v for row in ood for v in row], ncols), comp.root, h1z, h2z, fri_roots, last,
                      query_list, openings, nonce)


# ------------------------------------------------------------------------------ verifier
def verify(air_cls, proof, pub_inputs, options):
    """src/starks/verifier.rs:559-657, on oracle primitives only."""
    if len(proof.query_list) < options.fri_number_of_queries:
        return False
    n, blowup, offset = proof.trace_length, options.blowup_factor, options.coset_offset
    air = air_cls(n, pub_inputs, options)
    m = n * blowup
    w, g, lde_dom, roots = _domain(n, blowup, offset)
    fr = proof.trace_ood_frame_evaluations
    ncols = air.trace_columns
    f = lambda tr: O.lw_to_int(tr.to_field())
    # step 1
    t = O.Transcript()
    t.append(proof.lde_trace_merkle_roots[0])
    bcs = air.boundary_constraints()
    b_alpha = [f(t) for _ in bcs]
    b_beta = [f(t) for _ in bcs]
    t_alpha = [f(t) for _ in range(air.num_transition_constraints)]
    t_beta = [f(t) for _ in range(air.num_transition_constraints)]
    t.append(proof.composition_poly_root)
    while True:
        z = f(t)
        if z not in lde_dom and z not in roots:
            break
    t.append(_felt_bytes(proof.composition_poly_even_ood_evaluation))
    t.append(_felt_bytes(proof.composition_poly_odd_ood_evaluation))
    for i in range(fr.num_rows()):
        for v in fr.row(i):
            t.append(_felt_bytes(v))
    gamma_even, gamma_odd = f(t), f(t)
    coeffs = [[f(t) for _ in air.transition_offsets] for _ in range(ncols)]
    zetas = []
    for root in proof.fri_layers_merkle_roots:
        t.append(root)
        zetas.append(f(t))
    t.append(_felt_bytes(proof.fri_last_value))
    if O.grinding_zeros(t.challenge(), proof.nonce) < options.grinding_factor:
        return False
    t.append(proof.nonce.to_bytes(8, "big"))
    iotas = [t.to_usize() % m for _ in range(options.fri_number_of_queries)]
    # step 2
    bound = air.composition_poly_degree_bound()
    bz = pow(z, bound - n, P)
    boundary = 0
    for (col, step, value), a, b in zip(bcs, b_alpha, b_beta):
        boundary += (fr.row(0)[col] - value) * pow((z - pow(g, step, P)) % P, -1, P) * ((a * bz + b) % P)
    tr = air.compute_transition([fr.row(i) for i in range(fr.num_rows())])
    denom = pow((pow(z, n, P) - 1) % P, -1, P)
    last_root = roots[-1]
    ex = []
    for index in range(1, max(air.transition_exemptions) + 1):
        poly = [1]
        for k in range(1, index + 1):
            poly = poly_mul_linear(poly, pow(last_root, k, P))
        ex.append(poly_eval(poly, z))
    trans = 0
    for ev, deg, e, a, b in zip(tr, air.transition_degrees, air.transition_exemptions, t_alpha, t_beta):
        trans += denom * ev * ((a * pow(z, bound - n * (deg - 1), P) + b) % P) * (ex[e - 1] if e else 1)
    claimed = (proof.composition_poly_even_ood_evaluation + z * proof.composition_poly_odd_ood_evaluation) % P
    if claimed != (boundary + trans) % P:
        return False
    # step 3
    inv2 = pow(2, -1, P)
    for q, iota in zip(proof.query_list, iotas):
        x_inv = pow(lde_dom[iota], -1, P)
        v = q.layers_evaluations[0]
        for k, root in enumerate(proof.fri_layers_merkle_roots):
            size = m >> k
            isym = (iota + size // 2) % size
            ok = O.merkle_verify(root, iota, O.int_to_lw(q.layers_evaluations[k]), q.layers_auth_paths[k])
            ok &= O.merkle_verify(root, isym, O.int_to_lw(q.layers_evaluations_sym[k]), q.layers_auth_paths_sym[k])
            s = q.layers_evaluations_sym[k]
            v = ((v + s) * inv2 + zetas[k] * (v - s) * inv2 * x_inv) % P
            x_inv = x_inv * x_inv % P
            nxt = q.layers_evaluations[k + 1] if k + 1 < len(q.layers_evaluations) else proof.fri_last_value
            if not ok or v != nxt:
                return False
    # step 4
    for i, (iota, op) in enumerate(zip(iotas, proof.deep_poly_openings)):
        comp = np.stack([O.int_to_lw(op.lde_composition_poly_even_evaluation), O.int_to_lw(op.lde_composition_poly_odd_evaluation)])
        if not O.merkle_verify(proof.composition_poly_root, iota, comp, op.lde_composition_poly_proof):
            return False
        # (the reference computes the trace openings' Merkle check and discards it, verifier.rs:411-422)
        x = lde_dom[iota]
        dinv = pow((x - z * z) % P, -1, P)
        divs = [pow((x - z * pow(g, r, P)) % P, -1, P) for r in range(fr.num_rows())]
        acc = 0
        for col in range(fr.row_width):
            for r in range(fr.num_rows()):
                acc += (op.lde_trace_evaluations[col] - fr.row(r)[col]) * divs[r] * coeffs[col][r]
        acc += (op.lde_composition_poly_even_evaluation - proof.composition_poly_even_ood_evaluation) * dinv * gamma_even
        acc += (op.lde_composition_poly_odd_evaluation - proof.composition_poly_odd_ood_evaluation) * dinv * gamma_odd
        if acc % P != proof.query_list[i].layers_evaluations[0]:
            return False
    return True
