"""Parity of the CUDA path (through the C ABI) with the CPU oracle: bit-exact on every output."""
import os
import random

import numpy as np
import pytest

import lambdaworks_cairo_prover_b200 as P
from lambdaworks_cairo_prover_b200 import _native as N
from lambdaworks_cairo_prover_b200 import felt
from oracle import pyoracle as O
from util import edge_felts, random_felts

pytestmark = pytest.mark.gpu
MOD = felt.MODULUS


@pytest.fixture(scope="module")
def ctx():
    c = P.Context(0)
    yield c
    c.close()


def binop(ctx, op, a, b):
    out = np.empty_like(a)
    ctx.check(N.lib().s252_fe_binop(ctx.handle, op, N.ptr(a), N.ptr(b), N.ptr(out), a.shape[0], N.HOST))
    return out


# ---------------------------------------------------------------- field and hash primitives
def test_field_ops_bit_exact(ctx):
    e = edge_felts()
    r = random_felts(1, 500)
    a = np.concatenate([np.repeat(e, len(e), axis=0), r])
    b = np.concatenate([np.tile(e, (len(e), 1)), random_felts(2, 500)])
    ai, bi = felt.to_ints(a), felt.to_ints(b)
    for op, fn in ((0, lambda x, y: x * y % MOD), (1, lambda x, y: (x + y) % MOD), (2, lambda x, y: (x - y) % MOD)):
        got = binop(ctx, op, a, b)
        want = felt.from_ints([fn(x, y) for x, y in zip(ai, bi)])
        assert (got == want).all()
    nz = np.array([i for i, v in enumerate(ai) if v != 0][:200])
    got = binop(ctx, 3, a[nz], a[nz])
    assert felt.to_ints(got) == [pow(ai[i], -1, MOD) for i in nz]


def test_keccak256_on_device(ctx):
    rng = np.random.default_rng(5)
    for length in (0, 1, 31, 32, 40, 64, 135, 136, 137, 271, 272, 273, 1088, 1089):
        n = 33
        msgs = rng.integers(0, 256, size=(n, max(length, 1)), dtype=np.uint8)[:, :length].copy()
        out = np.empty((n, 32), dtype=np.uint8)
        flat = np.ascontiguousarray(msgs).reshape(-1)
        ctx.check(N.lib().s252_keccak256_batch(ctx.handle, N.ptr(flat) if length else None, length, n, N.ptr(out)))
        for i in range(n):
            assert out[i].tobytes() == O.keccak256(msgs[i].tobytes()), length


# ---------------------------------------------------------------- FFTPoly
@pytest.mark.parametrize("logn", list(range(0, 15)) + [16])
def test_interpolate_fft(ctx, logn):
    ev = random_felts(100 + logn, 1 << logn)
    got = np.empty_like(ev)
    ctx.check(N.lib().s252_interpolate_fft(ctx.handle, N.ptr(ev), ev.shape[0], N.ptr(got), N.HOST))
    assert (got == O.interpolate_fft(ev)).all()


def test_interpolate_fft_rejects_non_power_of_two(ctx):
    with pytest.raises(P.FFTError):
        P.Polynomial.interpolate_fft(random_felts(1, 12), ctx)


@pytest.mark.parametrize("logn", [0, 3, 11, 12, 13])
def test_interpolate_offset_fft(ctx, logn):
    ev = random_felts(200 + logn, 1 << logn)
    off = felt.from_int(3)
    got = np.empty_like(ev)
    ctx.check(N.lib().s252_interpolate_offset_fft(ctx.handle, N.ptr(ev), ev.shape[0], N.ptr(off), N.ptr(got), N.HOST))
    assert (got == O.interpolate_offset_fft(ev, off)).all()


@pytest.mark.parametrize("n_coeffs,blowup,domain", [
    (1, 1, None), (1, 4, 8), (5, 2, None), (8, 2, 8), (8, 4, 8), (9, 4, 8), (64, 8, 64), (100, 4, 128),
    (1 << 11, 4, 1 << 11), (1 << 12, 2, 1 << 12), (1 << 13, 4, 1 << 13), (3000, 4, 1 << 12), (1 << 10, 1, 1 << 14),
    (7, 1, 1 << 13), (1 << 14, 8, None)])
def test_evaluate_offset_fft(ctx, n_coeffs, blowup, domain):
    c = random_felts(300 + n_coeffs, n_coeffs)
    off = felt.from_int(3)
    got = P.Polynomial(c).evaluate_offset_fft(blowup, domain, off, ctx)
    want = O.evaluate_offset_fft(c, blowup, domain, off)
    assert got.shape == want.shape and (got == want).all()


def test_evaluate_offset_fft_zero_and_trimmed_polynomials(ctx):
    off = felt.from_int(7)
    z = P.Polynomial(np.zeros((4, 4), dtype=np.uint64))
    assert z.coeff_len() == 0
    got = z.evaluate_offset_fft(2, 4, off, ctx)
    assert got.shape[0] == 8 and not got.any()
    c = np.concatenate([random_felts(9, 5), np.zeros((3, 4), dtype=np.uint64)])
    got = P.Polynomial(c).evaluate_offset_fft(2, None, off, ctx)
    assert (got == O.evaluate_offset_fft(c, 2, None, off)).all() and got.shape[0] == 16


def test_lde_property_out_i_equals_p_at_h_w_i(ctx):
    # src/starks/prover.rs:838-862 against naive evaluation with python integers
    fib = [1, 1]
    while len(fib) < 8:
        fib.append(fib[-1] + fib[-2])
    poly = P.Polynomial.interpolate_fft(felt.from_ints(fib), ctx)
    out = P.evaluate_polynomial_on_lde_domain(poly, 2, 8, felt.from_int(3), ctx)
    w = O.lw_to_int(O.primitive_root(4))
    c = felt.to_ints(poly.coefficients)
    want = [sum(ci * pow(3 * pow(w, i, MOD), k, MOD) for k, ci in enumerate(c)) % MOD for i in range(16)]
    assert felt.to_ints(out) == want


def test_lde_edge_case_degree_ge_domain(ctx):
    # src/starks/prover.rs:864-881: x^8 on a domain of 8 with blowup 4 -> step rule
    coeffs = felt.from_ints([0] * 8 + [1])
    out = P.evaluate_polynomial_on_lde_domain(P.Polynomial(coeffs), 4, 8, felt.from_int(3), ctx)
    assert out.shape[0] == 32
    assert (out == O.evaluate_polynomial_on_lde_domain(coeffs, 4, 8, felt.from_int(3))).all()


def test_three_pass_transform_forced_small(ctx):
    """Forces the three-pass decomposition (used beyond 2^22) at small sizes."""
    old = os.environ.get("S252_MAX_LOGL")
    os.environ["S252_MAX_LOGL"] = "4"
    try:
        c3 = P.Context(0)
    finally:
        if old is None:
            del os.environ["S252_MAX_LOGL"]
        else:
            os.environ["S252_MAX_LOGL"] = old
    try:
        for logn in (5, 8, 9, 10, 11, 12):
            ev = random_felts(400 + logn, 1 << logn)
            got = np.empty_like(ev)
            c3.check(N.lib().s252_interpolate_fft(c3.handle, N.ptr(ev), ev.shape[0], N.ptr(got), N.HOST))
            assert (got == O.interpolate_fft(ev)).all(), logn
            got = P.Polynomial(ev).evaluate_offset_fft(4, None, felt.from_int(3), c3)
            assert (got == O.evaluate_offset_fft(ev, 4, None, felt.from_int(3))).all(), logn
    finally:
        c3.close()


def _shared_transform(c, log_n, inverse, blowup, parts, data):
    """s252_ntt_shared with every part played in turn by this one GPU: all phase-0 slabs, then all phase-1 row ranges.  The parts
    write disjoint positions of z / out, so the result must be the whole transform."""
    import ctypes as C
    L = N.lib()
    n = 1 << log_n
    cos = 1 if inverse else blowup
    d_in, d_z, d_out = c.device_alloc(n * 32), c.device_alloc(cos * n * 32), c.device_alloc(cos * n * 32)
    c.to_device(d_in, data)
    c.check(L.s252_convert_elements(c.handle, C.c_void_p(d_in), C.c_void_p(d_in), n, 1))
    l1 = C.c_uint()
    c.check(L.s252_ntt_shared(c.handle, log_n, int(inverse), cos, 3, 2, 0, parts, None, None, None, C.byref(l1)))
    for phase in (0, 1):
        for part in range(parts):
            c.check(L.s252_ntt_shared(c.handle, log_n, int(inverse), cos, 3, phase, part, parts, C.c_void_p(d_in), C.c_void_p(d_z),
                                      C.c_void_p(d_out), None))
    c.check(L.s252_convert_elements(c.handle, C.c_void_p(d_out), C.c_void_p(d_out), cos * n, 0))
    out = np.empty((cos * n, 4), dtype=np.uint64)
    c.to_host(out, d_out)
    for p_ in (d_in, d_z, d_out):
        c.device_free(p_)
    return out, int(l1.value)


@pytest.mark.parametrize("log_n,blowup,parts", [(12, 4, 2), (13, 2, 4), (16, 4, 8), (18, 8, 8)])
def test_shared_transform_two_pass(ctx, log_n, blowup, parts):
    """the four-step split of ONE column's transform over `parts` GPUs (here: one GPU playing every part)."""
    ev = random_felts(4100 + log_n, 1 << log_n)
    got, l1 = _shared_transform(ctx, log_n, True, 1, parts, ev)
    coeffs = O.interpolate_fft(ev)
    assert (got == coeffs).all()
    got, _ = _shared_transform(ctx, log_n, False, blowup, parts, coeffs)
    assert (got == O.evaluate_offset_fft(coeffs, blowup, 1 << log_n, felt.from_int(3))).all()
    assert 0 < l1 < log_n


def test_shared_transform_three_pass_forced_small():
    old = os.environ.get("S252_MAX_LOGL")
    os.environ["S252_MAX_LOGL"] = "5"
    try:
        c3 = P.Context(0)
    finally:
        if old is None:
            del os.environ["S252_MAX_LOGL"]
        else:
            os.environ["S252_MAX_LOGL"] = old
    try:
        for log_n, parts in ((11, 2), (13, 4), (15, 8)):
            ev = random_felts(4200 + log_n, 1 << log_n)
            got, _ = _shared_transform(c3, log_n, True, 1, parts, ev)
            coeffs = O.interpolate_fft(ev)
            assert (got == coeffs).all(), log_n
            got, _ = _shared_transform(c3, log_n, False, 4, parts, coeffs)
            assert (got == O.evaluate_offset_fft(coeffs, 4, 1 << log_n, felt.from_int(3))).all(), log_n
    finally:
        c3.close()


# ---------------------------------------------------------------- Merkle
@pytest.mark.parametrize("n,c", [(1, 1), (1, 5), (2, 1), (4, 2), (8, 17), (16, 18), (64, 33), (256, 34), (512, 35),
                                 (1024, 1), (2048, 2), (1 << 13, 4), (1 << 10, 52), (1 << 14, 1), (128, 16), (32, 68), (32, 69)])
def test_merkle_build(ctx, n, c):
    rows = random_felts(500 + n + c, n * c).reshape(n, c, 4)
    tree = P.BatchedMerkleTree.build(rows, ctx)
    want = O.merkle_build(rows)
    assert (tree.nodes() == want).all()
    assert tree.root == want[0].tobytes()
    rng = random.Random(n)
    for pos in {0, n - 1, rng.randrange(n)}:
        proof = tree.get_proof_by_pos(pos)
        assert [bytes(x) for x in O.merkle_path(want, pos)] == proof.merkle_path
        assert O.merkle_verify(tree.root, pos, rows[pos], proof.merkle_path)
    assert tree.get_proof_by_pos(n) is None
    tree.free()


def test_merkle_build_rejects_non_power_of_two(ctx):
    with pytest.raises(P.Stark252Error):
        P.BatchedMerkleTree.build(random_felts(1, 12).reshape(6, 2, 4), ctx)


def test_merkle_edge_values(ctx):
    e = edge_felts()
    rows = np.concatenate([e, e[:5]])[:16].reshape(8, 2, 4)
    tree = P.BatchedMerkleTree.build(rows, ctx)
    assert (tree.nodes() == O.merkle_build(rows)).all()


# ---------------------------------------------------------------- interpolate_and_commit
@pytest.mark.parametrize("logn,c,blowup", [(1, 1, 2), (3, 2, 4), (4, 3, 4), (8, 34, 4), (10, 18, 4), (11, 5, 8),
                                           (12, 3, 4), (13, 34, 4), (14, 2, 2), (15, 1, 8), (12, 33, 8)])
def test_interpolate_and_commit(ctx, logn, c, blowup):
    n = 1 << logn
    trace = random_felts(600 + logn * 7 + c, n * c).reshape(n, c, 4)
    want = O.interpolate_and_commit(trace, blowup, 3, threads=8)
    t = P.DefaultTranscript()
    opts = P.ProofOptions(blowup, 3, 3, 1)
    commit, root = P.interpolate_and_commit(P.TraceTable(trace.reshape(-1, 4), c), P.Domain(n, opts), t, ctx)
    assert root == want["root"]
    for j in range(c):
        assert (commit.coefficients(j) == want["coeffs"][j]).all(), j
        assert (commit.lde_column(j) == want["lde"][j]).all(), j
    assert (commit.nodes() == want["nodes"]).all()
    # the root went into the transcript (prover.rs:151)
    t2 = O.Transcript()
    t2.append(want["root"])
    assert t.challenge() == t2.challenge()
    # openings (prover.rs:484-529)
    idx = [0, n * blowup - 1, (n * blowup) // 3]
    rows, paths = commit.open(idx)
    for q, i in enumerate(idx):
        assert (rows[q] == want["lde"][:, i]).all()
        assert (paths[q] == O.merkle_path(want["nodes"], i)).all()
    commit.free()


def test_interpolate_and_commit_edge_tables(ctx):
    n, c = 64, 3
    for fill in (0, MOD - 1):
        trace = np.tile(felt.from_int(fill), (n, c, 1))
        want = O.interpolate_and_commit(trace, 4, 3)
        commit, root = P.interpolate_and_commit(P.TraceTable(trace.reshape(-1, 4), c), P.Domain(n, P.ProofOptions(4, 3, 3, 1)),
                                                P.DefaultTranscript(), ctx)
        assert root == want["root"]
        assert (commit.lde_column(1) == want["lde"][1]).all()


def test_interpolate_and_commit_rejects_bad_shapes(ctx):
    trace = random_felts(1, 12 * 2).reshape(12, 2, 4)
    with pytest.raises(P.FFTError):
        P.interpolate_and_commit(P.TraceTable(trace.reshape(-1, 4), 2), P.Domain(12, P.ProofOptions(4, 3, 3, 1)),
                                 P.DefaultTranscript(), ctx)


@pytest.mark.parametrize("logn,blowup,ncoef", [(4, 4, 16), (10, 4, 1000), (12, 4, 4096), (13, 2, 5000)])
def test_round2_lde_and_commit(ctx, logn, blowup, ncoef):
    # src/starks/prover.rs:254-276: H1, H2 -> LDE -> rows of 2 -> batch_commit
    n = 1 << logn
    h1, h2 = random_felts(700 + logn, ncoef), random_felts(701 + logn, ncoef - 3)
    dom = P.Domain(n, P.ProofOptions(blowup, 3, 3, 1))
    commit, root = P.lde_and_commit([P.Polynomial(h1), P.Polynomial(h2)], dom, ctx)
    e1 = O.evaluate_polynomial_on_lde_domain(h1, blowup, n, felt.from_int(3))
    e2 = O.evaluate_polynomial_on_lde_domain(h2, blowup, n, felt.from_int(3))
    nodes, want_root = O.commit_columns(np.stack([e1, e2]))
    assert root == want_root
    assert (commit.lde_column(0) == e1).all() and (commit.lde_column(1) == e2).all()
    assert (commit.nodes() == nodes).all()


# ---------------------------------------------------------------- FRI
@pytest.mark.parametrize("logn,blowup,ncoef", [(1, 2, 2), (3, 4, 8), (6, 4, 64), (10, 4, 1 << 10), (12, 4, 4000),
                                               (13, 8, 1 << 13), (11, 4, 5)])
def test_fri_commit_phase(ctx, logn, blowup, ncoef):
    n = 1 << logn
    M = n * blowup
    p0 = random_felts(800 + logn, ncoef)
    h = felt.from_int(3)
    t_gpu, t_ref = P.DefaultTranscript(), O.Transcript()
    t_gpu.append(b"fri-seed")
    t_ref.append(b"fri-seed")
    last, layers = P.fri_commit_phase(logn, P.Polynomial(p0), t_gpu, h, M, ctx)
    want_last, want_roots, want_evals, want_nodes = O.fri_commit_phase(logn, p0, t_ref, h, M)
    assert (last == want_last).all()
    assert len(layers) == logn
    for k, layer in enumerate(layers):
        assert layer.domain_size == M >> k
        assert layer.root == want_roots[k].tobytes()
        assert (layer.evaluation == want_evals[k]).all(), k
        assert (layer.nodes() == want_nodes[k]).all(), k
    assert t_gpu.challenge() == t_ref.challenge()
    # query phase (fri/mod.rs:74-127)
    queries, iotas = P.fri_query_phase(3, M, layers, t_gpu)
    assert iotas == [t_ref.to_usize() % M for _ in range(3)]
    for q, iota in zip(queries, iotas):
        for k in range(logn):
            size = M >> k
            i, isym = iota % size, (iota + size // 2) % size
            assert (q.layers_evaluations[k] == want_evals[k][i]).all()
            assert (q.layers_evaluations_sym[k] == want_evals[k][isym]).all()
            assert q.layers_auth_paths[k].merkle_path == [bytes(x) for x in O.merkle_path(want_nodes[k], i)]
            assert q.layers_auth_paths_sym[k].merkle_path == [bytes(x) for x in O.merkle_path(want_nodes[k], isym)]
            # un-reduced iota verifies too (verifier.rs:508)
            assert O.merkle_verify(layers[k].root, iota, q.layers_evaluations[k], q.layers_auth_paths[k].merkle_path)
    layers.free()


@pytest.mark.parametrize("logn,blowup", [(3, 2), (8, 4), (15, 4), (17, 2)])
def test_fri_layer_by_layer_interface(ctx, logn, blowup):
    """fri_layer0 / fri_fold_commit for a caller that keeps its own transcript == fri_commit_phase.  The larger sizes cross
    the thresholds of the single-launch tree top (2^9 and 2^17 leaves) and of the tail kernel."""
    import ctypes as C
    from lambdaworks_cairo_prover_b200 import _native as N
    L = N.lib()
    n = 1 << logn
    M = n * blowup
    p0 = random_felts(900 + logn, n)
    h = felt.from_int(3)
    t_ref = O.Transcript()
    t_ref.append(b"layers")
    want_last, want_roots, want_evals, _ = O.fri_commit_phase(logn, p0, t_ref, h, M)
    t = P.DefaultTranscript()
    t.append(b"layers")
    fri, root, last = C.c_void_p(), np.empty(32, dtype=np.uint8), np.empty(4, dtype=np.uint64)
    ctx.check(L.s252_fri_layer0(ctx.handle, N.ptr(p0), n, N.ptr(h), M, N.HOST, C.byref(fri), N.ptr(root)))
    t.append(root.tobytes())
    assert root.tobytes() == want_roots[0].tobytes()
    for k in range(1, logn):
        zeta = P.transcript_to_field(t)
        ctx.check(L.s252_fri_fold_commit(fri, N.ptr(zeta), N.ptr(root)))
        assert root.tobytes() == want_roots[k].tobytes(), k
        t.append(root.tobytes())
    zeta = P.transcript_to_field(t)
    ctx.check(L.s252_fri_fold_last(fri, N.ptr(zeta), N.ptr(last)))
    assert (last == want_last).all()
    t.append(felt.to_bytes_be(last))
    assert t.challenge() == t_ref.challenge()
    assert L.s252_fri_n_layers(fri) == logn
    ev = np.empty((M >> (logn - 1), 4), dtype=np.uint64)
    ctx.check(L.s252_fri_read_layer(fri, logn - 1, 0, ev.shape[0], N.ptr(ev)))
    assert (ev == want_evals[logn - 1]).all()
    L.s252_fri_destroy(fri)


def test_grinding_split_over_parts_gives_the_minimum(ctx):
    """disjoint nonce ranges + min == the sequential search, for every way of splitting."""
    import ctypes as C
    from lambdaworks_cairo_prover_b200 import _native as N
    L = N.lib()
    rng = np.random.default_rng(23)
    for factor in (3, 14, 20):
        ch = rng.integers(0, 256, size=32, dtype=np.uint8)
        want = O.generate_nonce_with_grinding(ch.tobytes(), factor)
        for parts, window_log in ((1, 32), (2, 32), (8, 32), (8, 21), (4, 20)):
            base, best = 0, (1 << 64) - 1
            while best == (1 << 64) - 1:                      # windows until some part finds a nonce, as the orchestration does
                found = []
                for part in range(parts):
                    f = C.c_uint64()
                    ctx.check(L.s252_grind_round(ctx.handle, N.ptr(ch), factor, base, 0, part, parts, window_log, C.byref(f)))
                    found.append(int(f.value))
                best = min(found)
                base += 1 << window_log
            assert best == want, (factor, parts, window_log, found)


def test_fri_rejects_oversized_polynomial(ctx):
    with pytest.raises(P.Stark252Error):
        P.fri_commit_phase(2, P.Polynomial(random_felts(1, 32)), P.DefaultTranscript(), felt.from_int(3), 16, ctx)


# ---------------------------------------------------------------- grinding
def test_grinding_kat(ctx):
    # src/starks/grinding.rs:56-64
    ch = bytes([226, 27, 133, 168, 62, 203, 20, 59, 122, 230, 227, 33, 76, 44, 53, 150, 200, 45,
                136, 162, 249, 239, 142, 90, 204, 191, 45, 4, 53, 22, 103, 240])
    assert P.generate_nonce_with_grinding(ch, 10, ctx) == 33


def test_grinding_returns_the_minimal_nonce(ctx):
    rng = np.random.default_rng(17)
    for factor in (0, 1, 4, 8, 12, 16, 18):
        ch = rng.integers(0, 256, size=32, dtype=np.uint8).tobytes()
        got = P.generate_nonce_with_grinding(ch, factor, ctx)
        assert got == O.generate_nonce_with_grinding(ch, factor)
        assert O.grinding_zeros(ch, got) >= factor
    ch = rng.integers(0, 256, size=32, dtype=np.uint8).tobytes()
    want = O.generate_nonce_with_grinding(ch, 12)
    assert P.generate_nonce_with_grinding(ch, 12, ctx, limit=want) is None      # None <-> Option::None
    assert P.generate_nonce_with_grinding(ch, 12, ctx, limit=want + 1) == want


def test_golden_proof_grinding_and_fri_roots(ctx):
    """The reference's golden proof: replay its transcript with the PRODUCT's transcript and
    redo grinding on the GPU -> the stored (minimal) nonce."""
    from conftest import GOLDEN
    from oracle.proof_format import read_proof_file
    proof, _, _ = read_proof_file(os.path.join(GOLDEN, "reference_proofs", "fibonacci_70000.proof"))
    t = P.DefaultTranscript()
    t.append(proof.lde_trace_merkle_roots[0])
    P.batch_sample_challenges(3, t)
    t.append(proof.lde_trace_merkle_roots[1])
    P.batch_sample_challenges(2 * 8 + 2 * 49, t)
    t.append(proof.composition_poly_root)
    P.transcript_to_field(t)
    t.append(proof.composition_poly_even_ood_evaluation.to_bytes(32, "big"))
    t.append(proof.composition_poly_odd_ood_evaluation.to_bytes(32, "big"))
    for v in proof.trace_ood_frame_evaluations.data:
        t.append(v.to_bytes(32, "big"))
    P.batch_sample_challenges(2 + 104, t)
    for root in proof.fri_layers_merkle_roots:
        t.append(root)
        P.transcript_to_field(t)
    t.append(proof.fri_last_value.to_bytes(32, "big"))
    assert P.generate_nonce_with_grinding(t.challenge(), 1, ctx) == proof.nonce
    # the opened FRI leaves of the golden proof hash to the same digests on the GPU
    q = proof.query_list[0]
    vals = felt.from_ints(q.layers_evaluations).reshape(-1, 1, 4)
    n = 1 << (len(vals) - 1).bit_length()
    pad = np.concatenate([vals, np.zeros((n - len(vals), 1, 4), dtype=np.uint64)])
    tree = P.FriMerkleTree.build(pad, ctx)
    leaves = tree.nodes(n - 1, len(vals))
    for k in range(len(vals)):
        assert leaves[k].tobytes() == O.keccak256(q.layers_evaluations[k].to_bytes(32, "big"))
