"""Known-answer and property tests of the CPU oracle against the reference's own unit tests, and
against naive python big-integer arithmetic (the ground truth for the field)."""
import random

import numpy as np
import pytest

from oracle import pyoracle as O

P = O.P


def rand_ints(rng, n):
    return [rng.randrange(P) for _ in range(n)]


# ------------------------------------------------------------------ reference KATs
def test_grinding_kat():
    # src/starks/grinding.rs:56-64
    ch = bytes([226, 27, 133, 168, 62, 203, 20, 59, 122, 230, 227, 33, 76, 44, 53, 150, 200, 45,
                136, 162, 249, 239, 142, 90, 204, 191, 45, 4, 53, 22, 103, 240])
    assert O.generate_nonce_with_grinding(ch, 10) == 33
    assert O.grinding_zeros(ch, 33) >= 10
    assert all(O.grinding_zeros(ch, n) < 10 for n in range(33))


def test_field_kat():
    # src/cairo/air.rs:1412-1451: 34/3, 34/11, 1
    f = lambda v: O.fe_from_u64(v)
    num = O.fe_sub(f(10), O.fe_add(f(3), O.fe_mul(f(15), f(5))))
    den = O.fe_sub(f(10), O.fe_add(f(1), O.fe_mul(f(15), f(1))))
    p0 = O.fe_mul(num, O.fe_inv(den))
    assert O.lw_to_int(p0) == 0x2aaaaaaaaaaaab0555555555555555555555555555555555555555555555561
    num = O.fe_sub(f(10), O.fe_add(f(1), O.fe_mul(f(15), f(1))))
    den = O.fe_sub(f(10), O.fe_add(f(2), O.fe_mul(f(15), f(2))))
    p1 = O.fe_mul(p0, O.fe_mul(num, O.fe_inv(den)))
    assert O.lw_to_int(p1) == 0x1745d1745d174602e8ba2e8ba2e8ba2e8ba2e8ba2e8ba2e8ba2e8ba2e8ba2ec
    num = O.fe_sub(f(10), O.fe_add(f(2), O.fe_mul(f(15), f(2))))
    den = O.fe_sub(f(10), O.fe_add(f(3), O.fe_mul(f(15), f(5))))
    p2 = O.fe_mul(p1, O.fe_mul(num, O.fe_inv(den)))
    assert O.lw_to_int(p2) == 1


def test_mask_kats():
    # src/starks/transcript.rs:96-131
    r = bytes([248] + [0] * 30 + [32])
    assert O.lw_to_int(O.randomness_to_field(r)) == 32
    r = bytes([255, 0] * 16)
    assert O.lw_to_int(O.randomness_to_field(r)) == int("0700FF00FF00FF00" + "FF00FF00FF00FF00" * 3, 16)


def test_fold_kat_shape():
    # src/starks/fri/fri_functions.rs:38-63 (there over F_293; the integers stay below p here)
    p0 = O.ints_to_lw([3, 1, 2, 7, 3, 5])
    p1 = O.fold_polynomial(p0, O.fe_from_u64(4))
    assert O.lw_to_ints(p1) == [7, 30, 23]
    p2 = O.fold_polynomial(p1, O.fe_from_u64(3))
    assert O.lw_to_ints(p2) == [97, 23]
    p3 = O.fold_polynomial(p2, O.fe_from_u64(2))
    assert O.lw_to_ints(p3) == [143]


def test_keccak_vectors():
    # Keccak-256 (0x01 padding) published vectors
    assert O.keccak256(b"").hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"
    assert O.keccak256(b"abc").hex() == "4e03657aea45a94fc7d47ba826c8d667c0d1e6e33a64a036ec44f58fa12d6c45"
    # crosses the 136-byte rate boundary
    assert O.keccak256(b"a" * 136) == O.keccak256(b"a" * 136)
    assert O.keccak256(b"a" * 135) != O.keccak256(b"a" * 136)


# ------------------------------------------------------------------ field vs python ints
def test_field_ops_match_python_ints():
    rng = random.Random(1)
    edge = [0, 1, 2, P - 1, P - 2, 2**251, 2**192, (1 << 251) + 1]
    vals = edge + rand_ints(rng, 40)
    for a in vals:
        for b in vals[:12]:
            A, B = O.int_to_lw(a), O.int_to_lw(b)
            assert O.lw_to_int(O.fe_add(A, B)) == (a + b) % P
            assert O.lw_to_int(O.fe_sub(A, B)) == (a - b) % P
            assert O.lw_to_int(O.fe_mul(A, B)) == a * b % P
            # results are fully reduced Montgomery representatives
            m = O.fe_mul(A, B)
            assert int.from_bytes(b"".join(int(x).to_bytes(8, "big") for x in m), "big") < P
        if a:
            assert O.lw_to_int(O.fe_inv(O.int_to_lw(a))) == pow(a, -1, P)
        assert O.fe_to_bytes_be(O.int_to_lw(a)) == a.to_bytes(32, "big")
        assert O.lw_to_int(O.fe_from_bytes_be(a.to_bytes(32, "big"))) == a
    assert O.lw_to_int(O.fe_from_u64(2**64 - 1)) == 2**64 - 1
    # from_bytes_be reduces non-canonical input
    assert O.lw_to_int(O.fe_from_bytes_be((P + 5).to_bytes(32, "big"))) == 5


def test_roots_of_unity_and_domain():
    # src/starks/prover.rs:788-835 (test_domain_constructor)
    for order in (1, 3, 4, 10):
        w = O.lw_to_int(O.primitive_root(order))
        assert pow(w, 2**order, P) == 1 and pow(w, 2**(order - 1), P) == P - 1
    n, b, h = 8, 2, 3
    w16 = O.lw_to_int(O.primitive_root(4))
    assert O.lw_to_int(O.primitive_root(3)) == pow(w16, b, P)
    coset = O.lw_to_ints(O.coset_powers(4, n * b, O.fe_from_u64(h)))
    assert coset == [h * pow(w16, i, P) % P for i in range(n * b)]


# ------------------------------------------------------------------ FFTPoly semantics
def naive_eval(coeffs, x):
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % P
    return acc


@pytest.mark.parametrize("n", [1, 2, 8, 64])
def test_interpolate_fft_inverts_evaluation(n):
    rng = random.Random(n)
    coeffs = rand_ints(rng, n)
    g = O.lw_to_int(O.primitive_root(n.bit_length() - 1))
    evals = [naive_eval(coeffs, pow(g, i, P)) for i in range(n)]
    assert O.lw_to_ints(O.interpolate_fft(O.ints_to_lw(evals))) == coeffs
    h = 7
    evals_h = [naive_eval(coeffs, h * pow(g, i, P) % P) for i in range(n)]
    assert O.lw_to_ints(O.interpolate_offset_fft(O.ints_to_lw(evals_h), O.fe_from_u64(h))) == coeffs


def test_interpolate_fft_rejects_non_power_of_two():
    with pytest.raises(ValueError):
        O.interpolate_fft(O.ints_to_lw([1, 2, 3]))


def test_lde_on_trace_polys():
    # src/starks/prover.rs:838-862: fibonacci trace of 8 rows, blowup 2, offset 3
    fib = [1, 1]
    while len(fib) < 8:
        fib.append((fib[-1] + fib[-2]) % P)
    poly = O.interpolate_fft(O.ints_to_lw(fib))
    out = O.evaluate_polynomial_on_lde_domain(poly, 2, 8, O.fe_from_u64(3))
    assert out.shape[0] == 16
    w = O.lw_to_int(O.primitive_root(4))
    c = O.lw_to_ints(poly)
    assert O.lw_to_ints(out) == [naive_eval(c, 3 * pow(w, i, P) % P) for i in range(16)]


def test_lde_edge_case_degree_ge_domain():
    # src/starks/prover.rs:864-881: x^8, blowup 4, domain 8 -> evaluated on 64 points, step 2
    coeffs = [0] * 8 + [1]
    out = O.evaluate_polynomial_on_lde_domain(O.ints_to_lw(coeffs), 4, 8, O.fe_from_u64(3))
    assert out.shape[0] == 32
    w = O.lw_to_int(O.primitive_root(5))
    assert O.lw_to_ints(out) == [pow(3 * pow(w, i, P) % P, 8, P) for i in range(32)]
    full = O.evaluate_offset_fft(O.ints_to_lw(coeffs), 4, 8, O.fe_from_u64(3))
    assert full.shape[0] == 64


def test_evaluate_offset_fft_trims_and_pads():
    rng = random.Random(5)
    coeffs = rand_ints(rng, 5) + [0, 0, 0]          # trailing zeros are trimmed: coeff_len = 5
    out = O.evaluate_offset_fft(O.ints_to_lw(coeffs), 2, None, O.fe_from_u64(3))
    assert out.shape[0] == 16                        # next_pow2(5) * 2
    w = O.lw_to_int(O.primitive_root(4))
    assert O.lw_to_ints(out) == [naive_eval(coeffs, 3 * pow(w, i, P) % P) for i in range(16)]
    zero = O.evaluate_offset_fft(O.ints_to_lw([0, 0]), 2, 4, O.fe_from_u64(3))
    assert zero.shape[0] == 8 and not zero.any()


# ------------------------------------------------------------------ Merkle
def test_merkle_build_paths_verify():
    rng = random.Random(9)
    for n, c in ((1, 3), (2, 1), (8, 2), (16, 5), (4, 34)):
        rows = O.ints_to_lw(rand_ints(rng, n * c)).reshape(n, c, 4)
        nodes = O.merkle_build(rows)
        # leaf / node rule
        msg = b"".join(O.fe_to_bytes_be(rows[0, j]) for j in range(c))
        assert nodes[n - 1].tobytes() == O.keccak256(msg)
        if n > 1:
            assert nodes[0].tobytes() == O.keccak256(nodes[1].tobytes() + nodes[2].tobytes())
        for pos in range(n):
            path = O.merkle_path(nodes, pos)
            assert path.shape[0] == n.bit_length() - 1
            assert O.merkle_verify(nodes[0].tobytes(), pos, rows[pos], path)
            if n > 1:
                assert not O.merkle_verify(nodes[0].tobytes(), pos ^ 1, rows[pos], path)
        assert O.merkle_path(nodes, n) is None
    with pytest.raises(ValueError):
        O.merkle_build(O.ints_to_lw([1, 2, 3]).reshape(3, 1, 4))


def test_commit_columns_equals_row_build():
    rng = random.Random(10)
    n, c = 16, 3
    rows = O.ints_to_lw(rand_ints(rng, n * c)).reshape(n, c, 4)
    cols = np.ascontiguousarray(rows.transpose(1, 0, 2))
    nodes, root = O.commit_columns(cols)
    assert (nodes == O.merkle_build(rows)).all() and root == nodes[0].tobytes()


# ------------------------------------------------------------------ interpolate_and_commit / FRI
def test_interpolate_and_commit_composes_primitives():
    rng = random.Random(11)
    n, c, b, h = 16, 3, 4, 3
    trace = O.ints_to_lw(rand_ints(rng, n * c)).reshape(n, c, 4)
    r = O.interpolate_and_commit(trace, b, h, threads=2)
    for j in range(c):
        col = np.ascontiguousarray(trace[:, j])
        poly = O.interpolate_fft(col)
        assert (r["coeffs"][j] == poly).all()
        assert (r["lde"][j] == O.evaluate_polynomial_on_lde_domain(poly, b, n, O.fe_from_u64(h))).all()
    rows = np.ascontiguousarray(r["lde"].transpose(1, 0, 2))
    assert (r["nodes"] == O.merkle_build(rows)).all()
    assert r["root"] == r["nodes"][0].tobytes()


def test_fri_commit_phase_matches_evaluation_domain_folding():
    """The reference folds coefficients and re-evaluates (fri/mod.rs:33-54); the GPU path folds in
    the evaluation domain with the verifier's formula (verifier.rs:511-512).  Same values."""
    rng = random.Random(12)
    n, b, h = 16, 4, 3
    M = n * b
    p0 = rand_ints(rng, n)
    t = O.Transcript()
    t.append(b"seed")
    last, roots, evals, nodes = O.fri_commit_phase(n.bit_length() - 1, O.ints_to_lw(p0), t, O.fe_from_u64(h), M)
    # independent replay with python ints
    t2 = O.Transcript()
    t2.append(b"seed")
    w = O.lw_to_int(O.primitive_root(M.bit_length() - 1))
    layer = [naive_eval(p0, h * pow(w, i, P) % P) for i in range(M)]
    off, size, inv2 = h, M, pow(2, -1, P)
    for k in range(n.bit_length() - 1):
        assert O.lw_to_ints(evals[k]) == layer
        tree = O.merkle_build(O.ints_to_lw(layer))
        assert (tree == nodes[k]).all() and roots[k].tobytes() == tree[0].tobytes()
        t2.append(tree[0].tobytes())
        zeta = O.lw_to_int(t2.to_field())
        wk = pow(w, 2**k, P)
        layer = [((layer[i] + layer[i + size // 2]) * inv2
                  + zeta * (layer[i] - layer[i + size // 2]) * inv2 * pow(off * pow(wk, i, P), -1, P)) % P
                 for i in range(size // 2)]
        off, size = off * off % P, size // 2
    assert len(set(layer)) == 1 and layer[0] == O.lw_to_int(last)
    t2.append(layer[0].to_bytes(32, "big"))
    assert t.challenge() == t2.challenge()
