"""The BASELINE-sized configurations against answers pinned offline with the CPU oracle
(tests/golden/synthetic/big_configs.json, written by tests/golden/make_big_configs.py): C3 single columns of 2^24 / 2^26 rows,
C4 (2^22 x 33, blowup 8), C5 (FRI from 2^24, 22 layers, grinding 20) and the Provable80Bits proof of fibonacci_70000.
No oracle runs on the GPU box."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from util import random_felts

PATH = os.path.join(GOLDEN, "synthetic", "big_configs.json")
CASES = json.load(open(PATH))["cases"] if os.path.exists(PATH) else {}


def need(name):
    if name not in CASES:
        pytest.skip("%s is not pinned in big_configs.json" % name)
    return CASES[name]


@pytest.fixture(scope="module")
def gpu():
    import lambdaworks_cairo_prover_b200 as P
    c = P.Context(0)
    yield P, c
    c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c3_20", "c3_24", "c3_26"])
def test_c3_single_column_root(gpu, name):
    P, ctx = gpu
    case = need(name)
    n = 1 << case["log_n"]
    trace = random_felts(case["seed"], n)
    commit, root = P.interpolate_and_commit(P.TraceTable(trace, 1), P.Domain(n, P.ProofOptions(case["blowup"], 3, case["offset"], 1)),
                                            P.DefaultTranscript(), ctx)
    commit.free()
    ctx.trim()
    assert root.hex() == case["root"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c4_small", "c4"])
def test_c4_root(gpu, name):
    P, ctx = gpu
    case = need(name)
    n, c = 1 << case["log_n"], case["cols"]
    trace = np.empty((n, c, 4), dtype=np.uint64)
    for j in range(c):
        trace[:, j, :] = random_felts(case["column_seed0"] + j, n)
    commit, root = P.interpolate_and_commit(P.TraceTable(trace.reshape(-1, 4), c), P.Domain(n, P.ProofOptions(case["blowup"], 3, case["offset"], 1)),
                                            P.DefaultTranscript(), ctx)
    commit.free()
    ctx.trim()
    assert root.hex() == case["root"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c5_small", "c5"])
def test_c5_fri_commit_phase(gpu, name):
    P, ctx = gpu
    from lambdaworks_cairo_prover_b200 import felt
    case = need(name)
    n = 1 << case["log_n"]
    p0 = random_felts(case["seed"], n)
    t = P.DefaultTranscript()
    t.append(bytes(32))
    last, layers = P.fri_commit_phase(case["log_n"], P.Polynomial(p0), t, felt.from_int(case["offset"]), n * case["blowup"], ctx)
    roots = [layer.root.hex() for layer in layers]
    layers.free()
    assert roots == case["roots"]
    assert felt.to_bytes_be(last).hex() == case["last_value"]
    assert P.generate_nonce_with_grinding(t.challenge(), case["grinding_factor"], ctx) == case["nonce"]
    ctx.trim()


@pytest.mark.gpu
def test_fib70000_provable80bits_proof_digest(gpu):
    P, ctx = gpu
    from lambdaworks_cairo_prover_b200 import cairo
    case = need("fib70000_80bits")
    regs, mem, size = cairo.run_program(cairo.fibonacci_program(case["fib_n"]))
    trace = cairo.build_main_trace(regs, mem, size)
    o = case["options"]
    proof = cairo.generate_cairo_proof(trace, P.ProofOptions(o["blowup_factor"], o["fri_number_of_queries"], o["coset_offset"], o["grinding_factor"]), ctx)
    assert len(proof) == case["proof_bytes"]
    assert hashlib.sha256(proof).hexdigest() == case["sha256"]
