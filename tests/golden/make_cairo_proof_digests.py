"""Writes tests/golden/cairo/proof_digests.json: sha256 of the serialized StarkProof the ORACLE produces for a few
fibonacci programs and option sets (CPU, a few seconds).  The GPU tests compare the product's proofs with these
digests without running the oracle on the GPU box."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from lambdaworks_cairo_prover_b200 import ProofOptions, cairo  # noqa: E402
from oracle.cairo_prover import cairo_prove  # noqa: E402

CASES = [(3, (4, 3, 3, 1)), (10, (4, 3, 3, 1)), (100, (8, 5, 7, 8)), (1000, (4, 31, 3, 10)), (2000, (2, 7, 5, 3))]


def main():
    out = []
    for n, o in CASES:
        regs, mem, size = cairo.run_program(cairo.fibonacci_program(n))
        t = cairo.build_main_trace(regs, mem, size)
        table = np.array(t.table).reshape(t.n_rows(), t.n_cols, 4)
        proof = cairo_prove(table, t.pub_inputs, ProofOptions(*o), threads=os.cpu_count()).serialize()
        out.append({"fib_n": n, "options": {"blowup_factor": o[0], "fri_number_of_queries": o[1], "coset_offset": o[2], "grinding_factor": o[3]},
                    "trace_rows": t.n_rows(), "proof_bytes": len(proof), "sha256": hashlib.sha256(proof).hexdigest()})
        print(out[-1])
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "cairo", "proof_digests.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
