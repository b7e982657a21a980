"""Pins the ORACLE's Cairo prover on the reference's own output: regenerates the fib(1,1,n) trace with
the library's Cairo machine, proves it with oracle/cairo_prover.py (CPU, test infrastructure) and
compares the serialized StarkProof with the reference's benches/proofs/fibonacci_70000.proof
(tests/golden/reference_proofs).  Takes ~2 minutes on 8 cores at n = 70000 (2^19 rows), so it is a
tool, not part of the CPU test suite; its result and the digests of every intermediate stage are
written to tests/golden/cairo/fib70000_stages.json, which the GPU tests use to localise a mismatch
without the slow oracle.

    python tools/cairo_golden_check.py            # fib(70000): byte-identical proof expected
"""
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from lambdaworks_cairo_prover_b200 import ProofOptions, cairo  # noqa: E402
from oracle.cairo_prover import cairo_prove  # noqa: E402
from oracle.proof_format import read_proof_file  # noqa: E402


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    n = 70000
    golden, golden_bytes, _ = read_proof_file(os.path.join(ROOT, "tests", "golden", "reference_proofs", "fibonacci_%d.proof" % n))
    regs, mem, size = cairo.run_program(cairo.fibonacci_program(n))
    trace = cairo.build_main_trace(regs, mem, size)
    table = np.array(trace.table).reshape(trace.n_rows(), trace.n_cols, 4)
    st = {}
    t0 = time.time()
    proof = cairo_prove(table, trace.pub_inputs, ProofOptions.default_test_options(), threads=os.cpu_count(), stages=st)
    secs = time.time() - t0
    same = proof.serialize() == golden_bytes
    out = {
        "program": "fibonacci_%d" % n, "trace_rows": trace.n_rows(), "oracle_seconds": round(secs, 1), "threads": os.cpu_count(),
        "proof_bytes_identical_to_reference": same,
        "main_root": st["main_root"].hex(), "aux_root": st["aux_root"].hex(), "composition_root": st["comp_root"].hex(),
        "sha256": {"main_trace_lw": digest(table), "rap_challenges_lw": digest(st["rap"]), "aux_trace_lw": digest(st["aux"]),
                   "constraint_evaluations_lw": digest(st["constraint_evals"]), "h1_coeffs_lw": digest(st["h1"]),
                   "h2_coeffs_lw": digest(st["h2"]), "deep_poly_coeffs_lw": digest(st["deep"])},
        "z": hex(st["z"]), "iotas": st["iotas"], "nonce": proof.nonce,
    }
    path = os.path.join(ROOT, "tests", "golden", "cairo", "fib70000_stages.json")
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out, indent=1))
    return 0 if same else 1


if __name__ == "__main__":
    sys.exit(main())
