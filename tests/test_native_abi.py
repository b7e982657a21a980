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
