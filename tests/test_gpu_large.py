"""Parity at BASELINE.json's sizes: the C2 main-trace commit (N = 2^19, 34 columns, blowup 4) bit-for-bit
against the oracle, the three-pass transform at its real size (2^23), a C5-shaped FRI commit
phase (2^22 -> 2^2), and size-independent properties at 2^24."""
import os

import numpy as np
import pytest

import lambdaworks_cairo_prover_b200 as P
from lambdaworks_cairo_prover_b200 import _native as N
from lambdaworks_cairo_prover_b200 import felt
from oracle import pyoracle as O
from util import random_felts

pytestmark = pytest.mark.gpu
MOD = felt.MODULUS


@pytest.fixture(scope="module")
def ctx():
    c = P.Context(0)
    yield c
    c.close()


def test_c2_main_trace_commit_full_size(ctx):
    n, c, blowup = 1 << 19, 34, 4
    trace = random_felts(0xB200 + 2, n * c).reshape(n, c, 4)
    t = P.DefaultTranscript()
    commit, root = P.interpolate_and_commit(P.TraceTable(trace.reshape(-1, 4), c), P.Domain(n, P.ProofOptions(blowup, 80, 3, 20)), t, ctx)
    want = O.interpolate_and_commit(trace, blowup, 3, threads=os.cpu_count() or 8, want_lde=True, want_nodes=True)
    assert root == want["root"]
    for j in (0, 17, 33):
        assert (commit.coefficients(j) == want["coeffs"][j]).all()
        assert (commit.lde_column(j) == want["lde"][j]).all()
    m = n * blowup
    assert (commit.nodes(0, 4096) == want["nodes"][:4096]).all()
    assert (commit.nodes(m - 1, 4096) == want["nodes"][m - 1:m - 1 + 4096]).all()
    idx = [0, 1, m - 1, 1066535 % m, 2059363 % m]
    rows, paths = commit.open(idx)
    for q, i in enumerate(idx):
        assert (rows[q] == want["lde"][:, i]).all()
        assert O.merkle_verify(root, i, rows[q], paths[q])
    commit.free()


def test_three_pass_transform_at_2_23(ctx):
    n = 1 << 23
    ev = random_felts(91, n)
    got = np.empty_like(ev)
    ctx.check(N.lib().s252_interpolate_fft(ctx.handle, N.ptr(ev), n, N.ptr(got), N.HOST))
    want = O.interpolate_fft(ev)
    assert (got == want).all()
    # and forward: evaluating the coefficients on the same domain returns the evaluations
    one = felt.from_int(1)
    back = P.Polynomial(got).evaluate_offset_fft(1, None, one, ctx)
    assert (back == ev).all()


def test_fri_commit_phase_c5_shape(ctx):
    logn, blowup = 20, 4
    n, m = 1 << logn, (1 << logn) * blowup
    p0 = random_felts(0xB200 + 5, n)
    h = felt.from_int(3)
    tg, tr = P.DefaultTranscript(), O.Transcript()
    tg.append(bytes(32))
    tr.append(bytes(32))
    last, layers = P.fri_commit_phase(logn, P.Polynomial(p0), tg, h, m, ctx)
    want_last, want_roots, _, _ = O.fri_commit_phase(logn, p0, tr, h, m, keep=False)
    assert (last == want_last).all()
    assert [layer.root for layer in layers] == [r.tobytes() for r in want_roots]
    ch = tg.challenge()
    assert ch == tr.challenge()
    assert P.generate_nonce_with_grinding(ch, 20, ctx) == O.generate_nonce_with_grinding(ch, 20)
    layers.free()


def test_single_column_lde_2_24_properties(ctx):
    """C3 shape beyond what the oracle does quickly: N = 2^22 coefficients, blowup 4 -> 2^24 points.
    Checked by (a) direct Horner evaluation at a few points, (b) coset interpolation round trip."""
    n, blowup = 1 << 22, 4
    coeffs = random_felts(77, n)
    off = felt.from_int(3)
    out = P.Polynomial(coeffs).evaluate_offset_fft(blowup, n, off, ctx)
    assert out.shape[0] == n * blowup
    w = O.lw_to_int(O.primitive_root(24))
    limbs = coeffs.astype(object)
    ints = [(int(a) << 192 | int(b) << 128 | int(c) << 64 | int(d)) for a, b, c, d in limbs]   # Montgomery residues
    rinv = pow(2**256, -1, MOD)
    for i in (0, 1, 12345677, n * blowup - 1):
        x = 3 * pow(w, i, MOD) % MOD
        acc = 0
        for cf in reversed(ints):
            acc = (acc * x + cf) % MOD
        assert felt.to_int(out[i]) == acc * rinv % MOD
    back = np.empty_like(out)
    ctx.check(N.lib().s252_interpolate_offset_fft(ctx.handle, N.ptr(out), out.shape[0], N.ptr(off), N.ptr(back), N.HOST))
    assert (back[:n] == coeffs).all() and not back[n:].any()
