"""Committed synthetic vectors (tests/golden/synthetic/commit_vectors.json): the oracle still
reproduces them (CPU), and the CUDA path matches them (GPU) without the oracle in the loop."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from util import random_felts

VECTORS = json.load(open(os.path.join(GOLDEN, "synthetic", "commit_vectors.json")))["cases"]
COMMITS = [c for c in VECTORS if c["kind"] == "interpolate_and_commit"]
FRIS = [c for c in VECTORS if c["kind"] == "fri_commit_phase"]


def test_oracle_reproduces_committed_vectors():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_synthetic", os.path.join(GOLDEN, "make_synthetic.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    c = COMMITS[1]
    assert m.commit_case(c["seed"], c["log_n"], c["cols"], c["blowup"], c["offset"]) == c
    f = FRIS[1]
    assert m.fri_case(f["seed"], f["log_n"], f["blowup"], f["offset"], f["grinding_factor"]) == f


@pytest.fixture(scope="module")
def gpu():
    import lambdaworks_cairo_prover_b200 as P
    c = P.Context(0)
    yield P, c
    c.close()


def keccak_hex(gpu, arr):
    """Digest through the product's own host Keccak (no oracle on this path)."""
    import ctypes as C
    from lambdaworks_cairo_prover_b200 import _native as N
    data = np.ascontiguousarray(arr).tobytes()
    out = (C.c_uint8 * 32)()
    N.lib().s252_keccak256((C.c_uint8 * len(data)).from_buffer_copy(data), len(data), out)
    return bytes(out).hex()


@pytest.mark.gpu
@pytest.mark.parametrize("case", COMMITS, ids=lambda c: "n%d_c%d_b%d" % (c["log_n"], c["cols"], c["blowup"]))
def test_gpu_commit_matches_committed_vectors(gpu, case):
    P, ctx = gpu
    n, c = 1 << case["log_n"], case["cols"]
    trace = random_felts(case["seed"], n * c)
    commit, root = P.interpolate_and_commit(P.TraceTable(trace, c), P.Domain(n, P.ProofOptions(case["blowup"], 3, case["offset"], 1)),
                                            P.DefaultTranscript(), ctx)
    assert root.hex() == case["root"]
    assert keccak_hex(gpu, np.stack([commit.coefficients(j) for j in range(c)])) == case["coeffs_digest"]
    assert keccak_hex(gpu, np.stack([commit.lde_column(j) for j in range(c)])) == case["lde_digest"]
    assert keccak_hex(gpu, commit.nodes()) == case["nodes_digest"]
    commit.free()


@pytest.mark.gpu
@pytest.mark.parametrize("case", FRIS, ids=lambda c: "n%d_b%d" % (c["log_n"], c["blowup"]))
def test_gpu_fri_matches_committed_vectors(gpu, case):
    P, ctx = gpu
    from lambdaworks_cairo_prover_b200 import felt
    n = 1 << case["log_n"]
    p0 = random_felts(case["seed"], n)
    t = P.DefaultTranscript()
    t.append(bytes(32))
    last, layers = P.fri_commit_phase(case["log_n"], P.Polynomial(p0), t, felt.from_int(case["offset"]), n * case["blowup"], ctx)
    assert [layer.root.hex() for layer in layers] == case["roots"]
    assert felt.to_bytes_be(last).hex() == case["last_value"]
    assert keccak_hex(gpu, np.concatenate([layer.evaluation for layer in layers])) == case["layers_digest"]
    assert P.generate_nonce_with_grinding(t.challenge(), case["grinding_factor"], ctx) == case["nonce"]
    layers.free()
