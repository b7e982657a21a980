"""Pins the CPU oracle on the reference's golden proof files.

Restates the verifier's transcript replay (src/starks/verifier.rs:59-206), FRI check (:443-523)
and DEEP/opening check (:358-441, :526-557) using ONLY oracle primitives, so a pass pins Keccak,
the transcript (byte reversal included), both Merkle back-ends, the field, the root-of-unity
constant, natural-order cosets, the fold formula, grinding and the wire format.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import pyoracle as O
from oracle.proof_format import StarkProof, read_proof_file

P = O.P
PROOFS = os.path.join(GOLDEN, "reference_proofs")

# (file, #boundary challenge pairs).  The 500/1000 files predate the change from one pair per
# column (52) to one pair per boundary constraint (8); SURVEY.md section 8c.
CASES = [("fibonacci_70000.proof", 8), ("fibonacci_500.proof", 52), ("fibonacci_1000.proof", 52)]
N_TRANSITION, N_RAP, N_COLS, N_MAIN, N_OFFSETS, QUERIES, GRINDING, BLOWUP, H = 49, 3, 52, 34, 2, 3, 1, 4, 3


def replay_transcript(proof, n_boundary):
    M = BLOWUP * proof.trace_length
    t = O.Transcript()
    t.append(proof.lde_trace_merkle_roots[0])                      # verifier.rs:78
    for _ in range(N_RAP):                                         # air.rs:731-737
        t.to_field()
    t.append(proof.lde_trace_merkle_roots[1])                      # verifier.rs:82-84
    for _ in range(2 * n_boundary + 2 * N_TRANSITION):             # verifier.rs:92-106
        t.to_field()
    t.append(proof.composition_poly_root)                          # verifier.rs:118
    z = O.lw_to_int(t.to_field())                                  # verifier.rs:125
    t.append(proof.composition_poly_even_ood_evaluation.to_bytes(32, "big"))
    t.append(proof.composition_poly_odd_ood_evaluation.to_bytes(32, "big"))
    fr = proof.trace_ood_frame_evaluations
    for i in range(fr.num_rows()):
        for e in fr.row(i):
            t.append(e.to_bytes(32, "big"))
    gamma_even = O.lw_to_int(t.to_field())
    gamma_odd = O.lw_to_int(t.to_field())
    coeffs = [[O.lw_to_int(t.to_field()) for _ in range(N_OFFSETS)] for _ in range(N_COLS)]
    zetas = []
    for root in proof.fri_layers_merkle_roots:                     # verifier.rs:165-174
        t.append(root)
        zetas.append(O.lw_to_int(t.to_field()))
    t.append(proof.fri_last_value.to_bytes(32, "big"))
    grinding_challenge = t.challenge()
    t.append(proof.nonce.to_bytes(8, "big"))
    iotas = [t.to_usize() % M for _ in range(QUERIES)]
    return dict(z=z, gamma_even=gamma_even, gamma_odd=gamma_odd, coeffs=coeffs, zetas=zetas,
                grinding_challenge=grinding_challenge, iotas=iotas)


@pytest.mark.parametrize("name,n_boundary", CASES)
def test_wire_format_round_trip(name, n_boundary):
    proof, proof_bytes, _ = read_proof_file(os.path.join(PROOFS, name))
    assert proof.serialize() == proof_bytes
    assert proof.trailing == b""
    assert len(proof.fri_layers_merkle_roots) == proof.trace_length.bit_length() - 1
    assert len(proof.lde_trace_merkle_roots) == 2
    # must-not-panic on truncation (fuzz/fuzz_targets/deserialize.rs)
    for cut in (0, 7, 8, 100, len(proof_bytes) - 1):
        with pytest.raises(ValueError):
            StarkProof.parse(proof_bytes[:cut])


@pytest.mark.parametrize("name,n_boundary", CASES)
def test_grinding_nonce_is_minimal_and_indices_open(name, n_boundary):
    proof, _, _ = read_proof_file(os.path.join(PROOFS, name))
    ch = replay_transcript(proof, n_boundary)
    assert O.grinding_zeros(ch["grinding_challenge"], proof.nonce) >= GRINDING
    assert O.generate_nonce_with_grinding(ch["grinding_challenge"], GRINDING) == proof.nonce
    # the sampled indices are the ones the stored Merkle paths open
    for q, iota in zip(proof.query_list, ch["iotas"]):
        assert O.merkle_verify(proof.fri_layers_merkle_roots[0], iota,
                               O.int_to_lw(q.layers_evaluations[0]), q.layers_auth_paths[0])


@pytest.mark.parametrize("name,n_boundary", CASES)
def test_fri_openings_and_folds(name, n_boundary):
    proof, _, _ = read_proof_file(os.path.join(PROOFS, name))
    ch = replay_transcript(proof, n_boundary)
    M = BLOWUP * proof.trace_length
    w = O.lw_to_int(O.primitive_root(M.bit_length() - 1))
    inv2 = pow(2, -1, P)
    for q, iota in zip(proof.query_list, ch["iotas"]):
        x_inv = pow(H * pow(w, iota, P) % P, -1, P)
        v = q.layers_evaluations[0]
        for k, root in enumerate(proof.fri_layers_merkle_roots):
            size = M >> k
            assert len(q.layers_auth_paths[k]) == size.bit_length() - 1
            isym = (iota + size // 2) % size
            # un-reduced iota on purpose: verifier.rs:508 passes it as is
            assert O.merkle_verify(root, iota, O.int_to_lw(q.layers_evaluations[k]), q.layers_auth_paths[k])
            assert O.merkle_verify(root, isym, O.int_to_lw(q.layers_evaluations_sym[k]), q.layers_auth_paths_sym[k])
            s = q.layers_evaluations_sym[k]
            v = ((v + s) * inv2 + ch["zetas"][k] * (v - s) * inv2 * x_inv) % P   # verifier.rs:511-512
            x_inv = x_inv * x_inv % P
            nxt = q.layers_evaluations[k + 1] if k + 1 < len(q.layers_evaluations) else proof.fri_last_value
            assert v == nxt


@pytest.mark.parametrize("name,n_boundary", CASES)
def test_deep_openings(name, n_boundary):
    proof, _, _ = read_proof_file(os.path.join(PROOFS, name))
    ch = replay_transcript(proof, n_boundary)
    M = BLOWUP * proof.trace_length
    w = O.lw_to_int(O.primitive_root(M.bit_length() - 1))
    g = pow(w, BLOWUP, P)
    assert g == O.lw_to_int(O.primitive_root(proof.trace_length.bit_length() - 1))   # prover.rs:821-824
    z, fr = ch["z"], proof.trace_ood_frame_evaluations
    for i, (iota, op) in enumerate(zip(ch["iotas"], proof.deep_poly_openings)):
        comp = np.stack([O.int_to_lw(op.lde_composition_poly_even_evaluation),
                         O.int_to_lw(op.lde_composition_poly_odd_evaluation)])
        assert O.merkle_verify(proof.composition_poly_root, iota, comp, op.lde_composition_poly_proof)
        main = O.ints_to_lw(op.lde_trace_evaluations[:N_MAIN])       # verifier.rs:403-408
        aux = O.ints_to_lw(op.lde_trace_evaluations[N_MAIN:])
        assert O.merkle_verify(proof.lde_trace_merkle_roots[0], iota, main, op.lde_trace_merkle_proofs[0])
        assert O.merkle_verify(proof.lde_trace_merkle_roots[1], iota, aux, op.lde_trace_merkle_proofs[1])
        x = H * pow(w, iota, P) % P
        denom_inv = pow((x - z * z) % P, -1, P)
        divs = [pow((x - z * pow(g, r, P)) % P, -1, P) for r in range(fr.num_rows())]
        acc = 0
        for col in range(fr.row_width):
            for r in range(fr.num_rows()):
                acc += (op.lde_trace_evaluations[col] - fr.row(r)[col]) * divs[r] * ch["coeffs"][col][r]
        acc += (op.lde_composition_poly_even_evaluation - proof.composition_poly_even_ood_evaluation) * denom_inv * ch["gamma_even"]
        acc += (op.lde_composition_poly_odd_evaluation - proof.composition_poly_odd_ood_evaluation) * denom_inv * ch["gamma_odd"]
        assert acc % P == proof.query_list[i].layers_evaluations[0]   # verifier.rs:436-437


def test_single_bit_mutations_are_rejected():
    proof, _, _ = read_proof_file(os.path.join(PROOFS, "fibonacci_500.proof"))
    ch = replay_transcript(proof, 52)
    q, iota = proof.query_list[0], ch["iotas"][0]
    root = proof.fri_layers_merkle_roots[0]
    good = O.int_to_lw(q.layers_evaluations[0])
    assert O.merkle_verify(root, iota, good, q.layers_auth_paths[0])
    assert not O.merkle_verify(root, iota ^ 1, good, q.layers_auth_paths[0])
    assert not O.merkle_verify(root, iota, O.int_to_lw(q.layers_evaluations[0] ^ 1), q.layers_auth_paths[0])
    bad_path = [bytes(q.layers_auth_paths[0][0][:-1] + bytes([q.layers_auth_paths[0][0][-1] ^ 1]))] + q.layers_auth_paths[0][1:]
    assert not O.merkle_verify(root, iota, good, bad_path)
    # a flipped root bit changes every later challenge
    proof.fri_layers_merkle_roots[0] = bytes([root[0] ^ 1]) + root[1:]
    assert replay_transcript(proof, 52)["iotas"] != ch["iotas"]
