"""CPU: the oracle's restatement of verify::<CairoAIR> (oracle/cairo_verifier.py) against the reference's
golden proof -- SURVEY.md section 8c's acceptance vector: verify(fibonacci_70000.proof) is true with the
public inputs stored in the file, and single-field mutations are rejected."""
import copy
import os

import numpy as np
import pytest

from conftest import GOLDEN
from lambdaworks_cairo_prover_b200 import ProofOptions, cairo, felt
from oracle import pyoracle as O
from oracle.cairo_prover import cairo_prove
from oracle.cairo_verifier import cairo_verify
from oracle.proof_format import StarkProof, read_proof_file
from test_cairo_trace import parse_public_inputs


class Pub:
    def __init__(self, d):
        self.pc_init, self.ap_init, self.fp_init, self.pc_final, self.ap_final = d["regs"]
        self.range_check_min, self.range_check_max, self.num_steps = d["rc_min"], d["rc_max"], d["num_steps"]
        self.public_memory = {a: O.int_to_lw(v) for a, v in d["public_memory"].items()}
        self.memory_segments = {}


def golden():
    proof, _, tail = read_proof_file(os.path.join(GOLDEN, "reference_proofs", "fibonacci_70000.proof"))
    return proof, Pub(parse_public_inputs(tail))


def test_reference_proof_is_accepted():
    proof, pub = golden()
    assert cairo_verify(proof, pub, ProofOptions.default_test_options())


def test_older_reference_proofs_pass_everything_but_the_composition_check():
    """fibonacci_500/1000 were written when 52+52 boundary coefficients were sampled: with that count the
    transcript, FRI and openings verify; only step 2 (whose formula changed) cannot."""
    for n in (500, 1000):
        proof, _, tail = read_proof_file(os.path.join(GOLDEN, "reference_proofs", "fibonacci_%d.proof" % n))
        pub = Pub(parse_public_inputs(tail))
        assert not cairo_verify(proof, pub, ProofOptions.default_test_options())
        assert not cairo_verify(proof, pub, ProofOptions.default_test_options(), n_boundary_pairs=52)


@pytest.mark.parametrize("mutate", [
    lambda p: setattr(p, "nonce", p.nonce + 1),
    lambda p: setattr(p, "fri_last_value", (p.fri_last_value + 1) % O.P),
    lambda p: setattr(p, "composition_poly_even_ood_evaluation", (p.composition_poly_even_ood_evaluation + 1) % O.P),
    lambda p: p.trace_ood_frame_evaluations.data.__setitem__(40, (p.trace_ood_frame_evaluations.data[40] + 1) % O.P),
    lambda p: p.trace_ood_frame_evaluations.data.__setitem__(52 + 19, (p.trace_ood_frame_evaluations.data[52 + 19] + 1) % O.P),
    lambda p: p.lde_trace_merkle_roots.__setitem__(1, bytes(32)),
    lambda p: p.fri_layers_merkle_roots.__setitem__(7, bytes([p.fri_layers_merkle_roots[7][0] ^ 1]) + p.fri_layers_merkle_roots[7][1:]),
    lambda p: p.query_list[1].layers_evaluations.__setitem__(3, (p.query_list[1].layers_evaluations[3] + 1) % O.P),
    lambda p: p.deep_poly_openings[2].lde_trace_evaluations.__setitem__(33, (p.deep_poly_openings[2].lde_trace_evaluations[33] + 1) % O.P),
    lambda p: setattr(p.deep_poly_openings[0], "lde_composition_poly_odd_evaluation", 5),
])
def test_mutated_reference_proof_is_rejected(mutate):
    proof, pub = golden()
    bad = copy.deepcopy(proof)
    mutate(bad)
    assert not cairo_verify(bad, pub, ProofOptions.default_test_options())


def test_wrong_public_inputs_are_rejected():
    proof, pub = golden()
    for field, delta in (("pc_final", 1), ("ap_final", 1), ("range_check_max", 1), ("num_steps", -1)):
        p2 = copy.deepcopy(pub)
        setattr(p2, field, getattr(p2, field) + delta)
        assert not cairo_verify(proof, p2, ProofOptions.default_test_options()), field
    p2 = copy.deepcopy(pub)
    p2.public_memory[6] = O.int_to_lw(69999)          # claims fib(69999) was run
    assert not cairo_verify(proof, p2, ProofOptions.default_test_options())


def test_oracle_proofs_verify():
    for n, opts in ((10, ProofOptions.default_test_options()), (30, ProofOptions(8, 7, 5, 6))):
        regs, mem, size = cairo.run_program(cairo.fibonacci_program(n))
        t = cairo.build_main_trace(regs, mem, size)
        table = np.array(t.table).reshape(t.n_rows(), t.n_cols, 4)
        proof = StarkProof.parse(cairo_prove(table, t.pub_inputs, opts, threads=2).serialize())
        assert cairo_verify(proof, t.pub_inputs, opts)
        assert not cairo_verify(proof, t.pub_inputs, ProofOptions(opts.blowup_factor, opts.fri_number_of_queries, opts.coset_offset,
                                                                   opts.grinding_factor + 12))
