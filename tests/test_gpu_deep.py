"""SURVEY section 8f, first "next" row: out-of-domain evaluations and the DEEP composition polynomial
on the GPU, against the oracle's coefficient-form restatement of prover.rs:410-482."""
import numpy as np
import pytest

import lambdaworks_cairo_prover_b200 as P
from lambdaworks_cairo_prover_b200 import felt
from oracle import pyoracle as O
from util import random_felts

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = P.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("logn,c_main,c_aux,blowup,offsets", [(3, 1, 0, 4, [0, 1, 2]), (6, 2, 1, 4, [0, 1]), (10, 5, 3, 4, [0, 1]),
                                                                (12, 34, 18, 4, [0, 1]), (9, 3, 0, 8, [0, 1, 2, 3]), (5, 2, 2, 2, [0])])
def test_round3_and_round4_on_device(ctx, logn, c_main, c_aux, blowup, offsets):
    n = 1 << logn
    m = n * blowup
    h = 3
    opts = P.ProofOptions(blowup, 3, h, 1)
    dom = P.Domain(n, opts)
    t_gpu, t_ref = P.DefaultTranscript(), O.Transcript()
    commits, polys = [], []
    for seed, cc in ((11, c_main), (12, c_aux)):
        if cc == 0:
            continue
        tr = random_felts(1000 * logn + seed, n * cc)
        cm, root = P.interpolate_and_commit(P.TraceTable(tr, cc), dom, t_gpu, ctx)
        t_ref.append(root)
        commits.append(cm)
        polys += [cm.coefficients(j) for j in range(cc)]
    polys = np.stack(polys)
    c = polys.shape[0]
    h12 = [random_felts(77 + logn, n), random_felts(78 + logn, n - 1)]
    comp, croot = P.lde_and_commit([P.Polynomial(x) for x in h12], dom, ctx)
    t_gpu.append(croot)
    t_ref.append(croot)
    z = P.transcript_to_field(t_gpu)
    assert (z == t_ref.to_field()).all()
    # round 3: OOD evaluations from the resident coefficients
    ood = P.get_trace_evaluations(commits, z, offsets, n, ctx)
    g = O.lw_to_int(O.primitive_root(logn))
    zi = felt.to_int(z)
    for k, off in enumerate(offsets):
        pt = felt.from_int(zi * pow(g, off, felt.MODULUS) % felt.MODULUS)
        for j in range(c):
            assert (ood[k, j] == O.poly_evaluate(polys[j], pt)).all(), (k, j)
    z2 = felt.from_int(zi * zi % felt.MODULUS)
    hz = P.evaluate_at(comp, z2)
    h1p = np.zeros((n, 4), dtype=np.uint64)
    h2p = np.zeros((n, 4), dtype=np.uint64)
    h1p[:len(h12[0])] = h12[0]
    h2p[:len(h12[1])] = h12[1]
    assert (hz[0] == O.poly_evaluate(h1p, z2)).all() and (hz[1] == O.poly_evaluate(h2p, z2)).all()
    # round 4
    gamma, gamma_p = P.transcript_to_field(t_gpu), P.transcript_to_field(t_gpu)
    gammas = np.stack(P.batch_sample_challenges(len(offsets) * c, t_gpu))
    for _ in range(2 + len(offsets) * c):
        t_ref.to_field()
    last, layers = P.fri_commit_phase_deep(logn, commits, comp, z, offsets, ood, hz[0], hz[1], gamma, gamma_p, gammas, t_gpu, h)
    p0 = O.deep_composition_poly(polys, h1p, h2p, z, offsets, ood, hz[0], hz[1], gamma, gamma_p, gammas.reshape(c, len(offsets), 4))
    want_last, want_roots, want_evals, _ = O.fri_commit_phase(logn, p0, t_ref, O.fe_from_u64(h), m)
    assert (layers[0].evaluation == want_evals[0]).all()          # p0 on the LDE coset, no NTT on the GPU side
    # the same polynomial block by block (what one rank of a sharded proof computes: s252_deep_rows builds its inverse tables
    # on the block + the halo the frame rotations reach back to; block 0's halo wraps around the end of the coset)
    import ctypes as C
    import torch
    from lambdaworks_cairo_prover_b200 import _native as N
    L = N.lib()
    offs = np.array(offsets, dtype=np.uint64)
    tabs = commits + [comp]
    for parts in (2, 8):
        rows = m // parts
        if rows == 0:
            continue
        for part in sorted({0, 1, parts // 2, parts - 1}):
            row0 = part * rows
            out = torch.empty((rows, 4), dtype=torch.int64, device="cuda:0")
            tables = (C.c_void_p * len(tabs))(*[L.s252_commit_device_lde(t.handle) + 32 * row0 for t in tabs])
            strides = (C.c_size_t * len(tabs))(*[m] * len(tabs))
            ncs = (C.c_size_t * len(tabs))(*[t.n_cols for t in tabs])
            ctx.check(L.s252_deep_rows(ctx.handle, tables, strides, ncs, len(tabs), row0, rows, m, n, N.ptr(z), N.ptr(offs), len(offsets),
                                       N.ptr(np.ascontiguousarray(ood.reshape(-1, 4))), N.ptr(np.ascontiguousarray(hz[0])),
                                       N.ptr(np.ascontiguousarray(hz[1])), N.ptr(gamma), N.ptr(gamma_p),
                                       N.ptr(np.ascontiguousarray(gammas.reshape(-1, 4))), h, C.c_void_p(out.data_ptr())))
            ctx.synchronize()
            # device buffers hold the library's internal layout: 8 x u32 least-significant first = the 4 u64 limbs of the
            # reference's layout in reverse order
            got = out.cpu().numpy().view(np.uint64)[:, ::-1]
            assert (got == want_evals[0][row0:row0 + rows]).all(), (parts, part)
    assert [layer.root for layer in layers] == [r.tobytes() for r in want_roots]
    assert (last == want_last).all()
    assert t_gpu.challenge() == t_ref.challenge()
    layers.free()
