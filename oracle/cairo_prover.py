"""ORACLE (test infrastructure, NOT the product): `generate_cairo_proof` = prove::<CairoAIR>
(src/cairo/air.rs:1183-1190, src/starks/prover.rs:532-776) restated on the CPU oracle's primitives.

Heavy loops are C (oracle/stark252_oracle.c, oracle/cairo_oracle.inc.c); this file is the round
structure and the Fiat-Shamir order.  PARITY: pinned -- for the regenerated fib(70000) trace the
proof returned here is byte-identical to the StarkProof section of the reference's
benches/proofs/fibonacci_70000.proof (tools/cairo_golden_check.py); the two older proof files pin the
round-1 roots.
"""
import numpy as np

from . import pyoracle as O
from .proof_format import DeepPolynomialOpenings, Frame, FriDecommitment, StarkProof

P = O.P
# column indices (src/cairo/air.rs:95-151)
MEM_P_TRACE_OFFSET, MEM_A_TRACE_OFFSET = 17, 19
RANGE_CHECK_COL_1, RANGE_CHECK_COL_3 = 43, 45
PERMUTATION_ARGUMENT_COL_3, PERMUTATION_ARGUMENT_RANGE_CHECK_COL_3 = 57, 60
BUILTIN_OFFSET = 9


def _i(a):
    return O.lw_to_int(a)


def boundary_constraints(pub, trace_length, rap, has_rc):
    """CairoAIR::boundary_constraints (air.rs:777-849) as (col, step, value LW) in the reference's order.
    pub: object with pc_init, ap_init, pc_final, ap_final, num_steps, range_check_min/max and
    public_memory {address: LW}."""
    bo = 0 if has_rc else BUILTIN_OFFSET
    alpha, z = _i(rap[0]), _i(rap[1])
    prod = 1
    for a, v in pub.public_memory.items():
        prod = prod * ((z - (a + alpha * _i(v))) % P) % P
    perm_final = pow(z, len(pub.public_memory), P) * pow(prod, -1, P) % P
    last = trace_length - 1
    f = O.int_to_lw
    return [(MEM_A_TRACE_OFFSET, 0, f(pub.pc_init)), (MEM_P_TRACE_OFFSET, 0, f(pub.ap_init)),
            (MEM_A_TRACE_OFFSET, pub.num_steps - 1, f(pub.pc_final)), (MEM_P_TRACE_OFFSET, pub.num_steps - 1, f(pub.ap_final)),
            (PERMUTATION_ARGUMENT_COL_3 - bo, last, f(perm_final)), (PERMUTATION_ARGUMENT_RANGE_CHECK_COL_3 - bo, last, f(1)),
            (RANGE_CHECK_COL_1 - bo, 0, f(pub.range_check_min)), (RANGE_CHECK_COL_3 - bo, last, f(pub.range_check_max))]


def sample_z_ood(t, n, m, offset):
    """src/starks/transcript.rs:53-70 -- membership tests by exponentiation instead of a scan."""
    hinv = pow(offset, -1, P)
    while True:
        z = _i(t.to_field())
        if pow(z * hinv % P, m, P) != 1 and pow(z, n, P) != 1:
            return z


def trim(c):
    c = np.asarray(c, dtype=np.uint64).reshape(-1, 4)
    n = c.shape[0]
    nz = np.nonzero(c.any(axis=1))[0]
    return c[: (nz[-1] + 1 if nz.size else 0)]


def cairo_prove(main_table, pub, options, threads=1, stages=None):
    """main_table: (n, c, 4) row-major LW.  Returns oracle.proof_format.StarkProof.
    stages: optional dict that receives intermediate products (for stage-by-stage parity tests)."""
    main_table = np.ascontiguousarray(main_table, dtype=np.uint64)
    n, c_main = main_table.shape[0], main_table.shape[1]
    has_rc = c_main > 34
    b, h = options.blowup_factor, options.coset_offset
    m = n * b
    order = n.bit_length() - 1
    g = _i(O.primitive_root(order))
    keep = stages if stages is not None else {}
    t = O.Transcript()
    # ---- round 1 (prover.rs:186-224)
    r_main = O.interpolate_and_commit(main_table, b, h, threads)
    t.append(r_main["root"])
    rap = np.stack([t.to_field() for _ in range(3)])                     # build_rap_challenges, air.rs:731-737
    addrs = sorted(pub.public_memory)
    aux = O.cairo_build_aux_trace(main_table, addrs, np.stack([pub.public_memory[a] for a in addrs]), rap)
    r_aux = O.interpolate_and_commit(aux, b, h, threads)
    t.append(r_aux["root"])
    polys = np.concatenate([r_main["coeffs"], r_aux["coeffs"]])
    lde = np.concatenate([r_main["lde"], r_aux["lde"]])
    ncols = polys.shape[0]
    keep.update(rap=rap, aux=aux, main_root=r_main["root"], aux_root=r_aux["root"])
    # ---- round 2 (prover.rs:598-640, 226-283)
    bcs = boundary_constraints(pub, n, rap, has_rc)
    nt = 50 if has_rc else 49
    b_alpha = [t.to_field() for _ in bcs]
    b_beta = [t.to_field() for _ in bcs]
    t_alpha = [t.to_field() for _ in range(nt)]
    t_beta = [t.to_field() for _ in range(nt)]
    bcoef = np.stack([np.stack([a, bb]) for a, bb in zip(b_alpha, b_beta)])
    tcoef = np.stack([np.stack([a, bb]) for a, bb in zip(t_alpha, t_beta)])
    evals = O.cairo_constraint_evaluations(lde, n, b, h, rap, bcs, bcoef, tcoef, has_rc, threads)
    hpoly = trim(O.interpolate_offset_fft(evals, O.fe_from_u64(h)))
    h1, h2 = trim(hpoly[0::2]), trim(hpoly[1::2])
    off = O.fe_from_u64(h)
    h1_lde = O.evaluate_polynomial_on_lde_domain(h1, b, n, off)
    h2_lde = O.evaluate_polynomial_on_lde_domain(h2, b, n, off)
    comp_nodes, comp_root = O.commit_columns(np.stack([h1_lde, h2_lde]))
    t.append(comp_root)
    keep.update(bcoef=bcoef, tcoef=tcoef, boundary=bcs, constraint_evals=evals, h1=h1, h2=h2, comp_root=comp_root)
    # ---- round 3 (prover.rs:650-690)
    z = sample_z_ood(t, n, m, h)
    z2 = z * z % P
    h1z = O.poly_evaluate(h1, O.int_to_lw(z2)) if len(h1) else O.int_to_lw(0)
    h2z = O.poly_evaluate(h2, O.int_to_lw(z2)) if len(h2) else O.int_to_lw(0)
    offsets = [0, 1]
    ood = np.stack([np.stack([O.poly_evaluate(polys[j], O.int_to_lw(z * pow(g, k, P) % P)) for j in range(ncols)]) for k in offsets])
    t.append(O.fe_to_bytes_be(h1z))
    t.append(O.fe_to_bytes_be(h2z))
    for row in ood:
        for v in row:
            t.append(O.fe_to_bytes_be(v))
    keep.update(z=z, ood=ood, h1z=h1z, h2z=h2z)
    # ---- round 4 (prover.rs:327-404)
    gamma, gamma_p = t.to_field(), t.to_field()
    tg = np.stack([t.to_field() for _ in range(len(offsets) * ncols)]).reshape(ncols, len(offsets), 4)
    pad = lambda p: np.concatenate([p, np.zeros((n - len(p), 4), dtype=np.uint64)]) if len(p) < n else p
    assert len(h1) <= n and len(h2) <= n, "composition polynomial exceeds the degree bound (invalid trace)"
    deep = O.deep_composition_poly(polys, pad(h1), pad(h2), O.int_to_lw(z), offsets, ood, h1z, h2z, gamma, gamma_p, tg)
    layers = order
    last, fri_roots, fevals, fnodes = O.fri_commit_phase(layers, trim(deep), t, off, m)
    nonce = O.generate_nonce_with_grinding(t.challenge(), options.grinding_factor)
    assert nonce is not None, "nonce not found"
    t.append(nonce.to_bytes(8, "big"))
    iotas = [t.to_usize() % m for _ in range(options.fri_number_of_queries)]
    path = lambda nodes, i: [bytes(p) for p in O.merkle_path(nodes, i)]
    query_list, openings = [], []
    for iota in iotas:                                                   # fri_query_phase, fri/mod.rs:74-127
        ev, ev_sym, pa, pa_sym = [], [], [], []
        for k in range(layers):
            size = m >> k
            i, isym = iota % size, (iota + size // 2) % size
            ev.append(_i(fevals[k][i]))
            ev_sym.append(_i(fevals[k][isym]))
            pa.append(path(fnodes[k], i))
            pa_sym.append(path(fnodes[k], isym))
        query_list.append(FriDecommitment(pa_sym, ev_sym, ev, pa))
    for iota in iotas:                                                   # open_deep_composition_poly, prover.rs:484-529
        idx = iota % m
        openings.append(DeepPolynomialOpenings(path(comp_nodes, idx), _i(h1_lde[idx]), _i(h2_lde[idx]),
                                               [path(r_main["nodes"], idx), path(r_aux["nodes"], idx)],
                                               [_i(lde[j][idx]) for j in range(ncols)]))
    keep.update(iotas=iotas, deep=deep)
    return StarkProof(n, [r_main["root"], r_aux["root"]], Frame([_i(v) for row in ood for v in row], ncols), comp_root,
                      _i(h1z), _i(h2z), [bytes(r) for r in fri_roots], _i(last), query_list, openings, nonce)
