"""CPU-side checks of the product library: it loads, exports every symbol the header declares,
its host-side transcript/field agree with the oracle, and it refuses to run without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from lambdaworks_cairo_prover_b200 import _native as N
import lambdaworks_cairo_prover_b200 as P
from oracle import pyoracle as O


def header_symbols():
    text = "".join(open(os.path.join(ROOT, "include", h)).read() for h in ("stark252_b200.h", "stark252_cairo.h"))
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(s252_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(N.library_path()) if os.path.exists(N.library_path()) else N.lib()
    syms = header_symbols()
    assert len(syms) >= 40
    for s in syms:
        assert hasattr(lib, s), "missing export " + s
    # and the python binding covers the same set
    assert sorted(N.SIGNATURES) == syms


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(P.Stark252Error):
        P.Context(0)
    with pytest.raises(P.Stark252Error):
        P.Polynomial.interpolate_fft(np.zeros((4, 4), dtype=np.uint64))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "lambdaworks_cairo_prover_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower() or f == "__init__.py" and False, f


def test_host_transcript_matches_oracle():
    a, b = P.DefaultTranscript(), O.Transcript()
    rng = np.random.default_rng(3)
    for i in range(40):
        data = rng.integers(0, 256, size=int(rng.integers(0, 300)), dtype=np.uint8).tobytes()
        a.append(data)
        b.append(data)
        if i % 3 == 0:
            assert a.challenge() == b.challenge()
        if i % 5 == 0:
            assert (P.transcript_to_field(a) == b.to_field()).all()
        if i % 7 == 0:
            assert P.transcript_to_usize(a) == b.to_usize()


def test_felt_conversions_match_oracle():
    from lambdaworks_cairo_prover_b200 import felt
    for v in (0, 1, 3, felt.MODULUS - 1, 2**200 + 17):
        assert (felt.from_int(v) == O.int_to_lw(v)).all()
        assert felt.to_int(O.int_to_lw(v)) == v
        assert felt.to_bytes_be(felt.from_int(v)) == O.fe_to_bytes_be(O.int_to_lw(v))


def test_evaluate_offset_fft_len_rule():
    L = N.lib()
    assert L.s252_evaluate_offset_fft_len(5, 2, 0) == 16
    assert L.s252_evaluate_offset_fft_len(9, 4, 8) == 64
    assert L.s252_evaluate_offset_fft_len(0, 2, 4) == 8
    assert L.s252_evaluate_offset_fft_len(8, 1, 64) == 64


@pytest.mark.parametrize("name", ["fibonacci_500", "fibonacci_70000"])
def test_c_serializer_reproduces_the_reference_proof_files(name):
    """s252_cairo_serialize_proof (host code; the sharded prover assembles its proof with it) must emit exactly
    StarkProof::serialize (proof/stark.rs:161-218): the pieces parsed out of the reference's own proof files go back in
    and the file's bytes must come out."""
    from conftest import GOLDEN
    from lambdaworks_cairo_prover_b200 import felt
    from oracle.proof_format import read_proof_file
    proof, raw, _ = read_proof_file(os.path.join(GOLDEN, "reference_proofs", name + ".proof"))
    L = N.lib()
    q, layers = len(proof.query_list), len(proof.fri_layers_merkle_roots)
    depth = len(proof.deep_poly_openings[0].lde_composition_poly_proof)
    cols = proof.trace_ood_frame_evaluations.row_width
    main_cols = 34
    f = lambda vals: felt.from_ints(list(vals))                                                   # noqa: E731
    ood, hz, last = f(proof.trace_ood_frame_evaluations.data), f([proof.composition_poly_even_ood_evaluation,
                                                                  proof.composition_poly_odd_ood_evaluation]), f([proof.fri_last_value])
    ev = np.stack([f(d.layers_evaluations) for d in proof.query_list])
    evs = np.stack([f(d.layers_evaluations_sym) for d in proof.query_list])

    def paths(get, n_layers):
        out = np.zeros((q, n_layers, depth, 32), dtype=np.uint8)
        for a in range(q):
            for k, p in enumerate(get(a)):
                assert len(p) == depth - k if n_layers > 1 else len(p) == depth
                out[a, k, :len(p)] = np.frombuffer(b"".join(p), dtype=np.uint8).reshape(-1, 32)
        return out
    pa = paths(lambda a: proof.query_list[a].layers_auth_paths, layers)
    pas = paths(lambda a: proof.query_list[a].layers_auth_paths_sym, layers)
    o = proof.deep_poly_openings
    comp_rows = np.stack([f([x.lde_composition_poly_even_evaluation, x.lde_composition_poly_odd_evaluation]) for x in o])
    comp_paths = paths(lambda a: [o[a].lde_composition_poly_proof], 1)[:, 0]
    rows = np.stack([f(x.lde_trace_evaluations) for x in o])
    main_rows, aux_rows = np.ascontiguousarray(rows[:, :main_cols]), np.ascontiguousarray(rows[:, main_cols:])
    main_paths = paths(lambda a: [o[a].lde_trace_merkle_proofs[0]], 1)[:, 0]
    aux_paths = paths(lambda a: [o[a].lde_trace_merkle_proofs[1]], 1)[:, 0]
    roots = [np.frombuffer(r, dtype=np.uint8) for r in proof.lde_trace_merkle_roots + [proof.composition_poly_root]]
    fr = np.frombuffer(b"".join(proof.fri_layers_merkle_roots), dtype=np.uint8)
    out, ln = C.c_void_p(), C.c_size_t()
    c = np.ascontiguousarray
    rc = L.s252_cairo_serialize_proof(proof.trace_length, N.ptr(roots[0]), N.ptr(roots[1]), N.ptr(roots[2]), N.ptr(c(ood)), cols, N.ptr(c(hz)), layers,
                                      N.ptr(fr), N.ptr(c(last)), q, depth, N.ptr(c(evs)), N.ptr(c(ev)), N.ptr(c(pas)), N.ptr(c(pa)), N.ptr(c(comp_rows)),
                                      N.ptr(c(comp_paths)), N.ptr(main_rows), main_cols, N.ptr(c(main_paths)), N.ptr(aux_rows), cols - main_cols,
                                      N.ptr(c(aux_paths)), proof.nonce, C.byref(out), C.byref(ln))
    assert rc == 0, L.s252_cairo_last_error()
    got = C.string_at(out.value, ln.value)
    L.s252_cairo_proof_free(out)
    assert got == raw
