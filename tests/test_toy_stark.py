"""prove -> verify for the reference's toy AIRs (tests/integration_tests.rs:36-112) on the restated
prover/verifier of tests/toy_stark.py; on the GPU the hot path runs through the C ABI and the
serialized StarkProof must be byte-identical to the oracle-backed one."""
import copy

import pytest

import toy_stark as TS
from lambdaworks_cairo_prover_b200.options import ProofOptions
from oracle.proof_format import StarkProof

CASES = [
    ("fib_8", TS.FibonacciAIR, lambda: TS.FibonacciAIR.trace(1, 1, 8), [1, 1], ProofOptions.default_test_options()),
    ("fib_64_b8", TS.FibonacciAIR, lambda: TS.FibonacciAIR.trace(1, 1, 64), [1, 1], ProofOptions(8, 5, 3, 6)),
    ("fib_2_cols_16", TS.Fibonacci2ColsAIR, lambda: TS.Fibonacci2ColsAIR.trace(1, 1, 16), [1, 1], ProofOptions.default_test_options()),
    ("quadratic_4", TS.QuadraticAIR, lambda: TS.QuadraticAIR.trace(3, 4), [3], ProofOptions.default_test_options()),
    ("quadratic_32", TS.QuadraticAIR, lambda: TS.QuadraticAIR.trace(3, 32), [3], ProofOptions(4, 4, 7, 4)),
]


@pytest.mark.parametrize("name,air,trace,pub,opts", CASES, ids=[c[0] for c in CASES])
def test_oracle_prove_then_verify(name, air, trace, pub, opts):
    proof = TS.prove(air, trace(), pub, opts, TS.OracleBackend())
    assert TS.verify(air, proof, pub, opts)
    # wire format round trip (src/starks/proof/stark.rs:596-707)
    data = proof.serialize()
    assert StarkProof.parse(data).serialize() == data
    assert TS.verify(air, StarkProof.parse(data), pub, opts)


def test_verifier_rejects_tampering():
    air, pub, opts = TS.FibonacciAIR, [1, 1], ProofOptions.default_test_options()
    proof = TS.prove(air, TS.FibonacciAIR.trace(1, 1, 8), pub, opts, TS.OracleBackend())
    assert TS.verify(air, proof, pub, opts)
    assert not TS.verify(air, proof, [1, 2], opts)                       # wrong public input
    for mutate in (lambda p: setattr(p, "fri_last_value", (p.fri_last_value + 1) % TS.P),
                   lambda p: setattr(p, "nonce", p.nonce + 1),
                   lambda p: setattr(p, "composition_poly_even_ood_evaluation", 5),
                   lambda p: p.fri_layers_merkle_roots.__setitem__(1, bytes(32)),
                   lambda p: p.query_list[0].layers_evaluations.__setitem__(0, 7),
                   lambda p: p.deep_poly_openings[0].lde_trace_evaluations.__setitem__(0, 9)):
        bad = copy.deepcopy(proof)
        mutate(bad)
        assert not TS.verify(air, bad, pub, opts)
    # an invalid trace does not verify (tests/integration_tests.rs negative cases)
    bad_trace = TS.FibonacciAIR.trace(1, 1, 8)
    bad_trace[0][5] += 1
    bad = TS.prove(air, bad_trace, pub, opts, TS.OracleBackend())
    assert not TS.verify(air, bad, pub, opts)
    # too few queries (src/starks/verifier.rs:570)
    assert not TS.verify(air, proof, pub, ProofOptions(4, 80, 3, 1))


@pytest.mark.gpu
@pytest.mark.parametrize("name,air,trace,pub,opts", CASES, ids=[c[0] for c in CASES])
def test_gpu_proof_bytes_equal_oracle_proof_bytes(name, air, trace, pub, opts):
    import lambdaworks_cairo_prover_b200 as P
    ctx = P.Context(0)
    try:
        got = TS.prove(air, trace(), pub, opts, TS.GpuBackend(ctx))
        want = TS.prove(air, trace(), pub, opts, TS.OracleBackend())
        assert got.serialize() == want.serialize()
        assert TS.verify(air, got, pub, opts)
        # rounds 3 and 4 on the GPU as well (OOD evaluations + DEEP polynomial in the evaluation domain)
        got2 = TS.prove(air, trace(), pub, opts, TS.GpuBackend(ctx, deep_on_device=True))
        assert got2.serialize() == want.serialize()
    finally:
        ctx.close()
