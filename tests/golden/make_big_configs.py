"""Pins the BASELINE-sized configurations with the CPU oracle (offline, minutes each):

    python tests/golden/make_big_configs.py [c5 c3_24 c4 c3_26 fib70000_80bits ...]

writes / updates tests/golden/synthetic/big_configs.json.  The GPU tests and bench.py compare the CUDA path
against this file (no oracle on the GPU box).  Inputs (the GPU side regenerates them from the same seeds):

  c4      SURVEY 8d C4: 2^22 rows x 33 columns, blowup 8, offset 3; COLUMN j = random_felts(C4_SEED + j, 2^22)
          (column-wise seeds so that a rank of a column-sharded commit generates only its own columns)
  c5      SURVEY 8d C5: fri_commit_phase(22 layers) of p0 = random_felts(0xB205, 2^22) over the 2^24-point coset of offset 3,
          transcript seeded with 32 zero bytes, then grinding at factor 20
  c3_LL   SURVEY 8d C3: one column of 2^LL rows = random_felts(0xB203 + LL, 2^LL), blowup 4, offset 3
  fib70000_80bits   sha256 of the oracle's proof of fibonacci_70000 under ProofOptions::new_secure(Provable80Bits, 3)
                    (benches/criterion_prover_70k.rs:48)
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import pyoracle as O   # noqa: E402
from util import random_felts      # noqa: E402

OUT = os.path.join(HERE, "synthetic", "big_configs.json")
C4_SEED = 0xB2040000
THREADS = os.cpu_count() or 1


def c4(log_n=22, cols=33, blowup=8, offset=3):
    n = 1 << log_n
    trace = np.empty((n, cols, 4), dtype=np.uint64)
    for j in range(cols):
        trace[:, j, :] = random_felts(C4_SEED + j, n)
    r = O.interpolate_and_commit(trace, blowup, offset, threads=THREADS, want_lde=False, want_nodes=False)
    return {"kind": "interpolate_and_commit", "column_seed0": C4_SEED, "log_n": log_n, "cols": cols, "blowup": blowup, "offset": offset,
            "root": r["root"].hex()}


def c3(log_n, blowup=4, offset=3):
    n = 1 << log_n
    seed = 0xB203 + log_n
    trace = random_felts(seed, n).reshape(n, 1, 4)
    r = O.interpolate_and_commit(trace, blowup, offset, threads=1, want_lde=False, want_nodes=False)
    return {"kind": "interpolate_and_commit", "seed": seed, "log_n": log_n, "cols": 1, "blowup": blowup, "offset": offset,
            "root": r["root"].hex()}


def c5(log_n=22, blowup=4, offset=3, grinding=20, seed=0xB205):
    n = 1 << log_n
    p0 = random_felts(seed, n)
    t = O.Transcript()
    t.append(bytes(32))
    last, roots, _, _ = O.fri_commit_phase(log_n, p0, t, O.fe_from_u64(offset), n * blowup, keep=False)
    ch = t.challenge()
    return {"kind": "fri_commit_phase", "seed": seed, "log_n": log_n, "blowup": blowup, "offset": offset,
            "roots": [x.tobytes().hex() for x in roots], "last_value": O.fe_to_bytes_be(last).hex(),
            "grinding_factor": grinding, "nonce": O.generate_nonce_with_grinding(ch, grinding)}


def fib_80bits(fib_n=70000):
    from lambdaworks_cairo_prover_b200 import ProofOptions, cairo
    from oracle.cairo_prover import cairo_prove
    regs, mem, size = cairo.run_program(cairo.fibonacci_program(fib_n))
    t = cairo.build_main_trace(regs, mem, size)
    table = np.array(t.table).reshape(t.n_rows(), t.n_cols, 4)
    opts = ProofOptions.new_secure("Provable80Bits", 3)
    proof = cairo_prove(table, t.pub_inputs, opts, threads=THREADS).serialize()
    return {"kind": "cairo_proof", "fib_n": fib_n, "trace_rows": t.n_rows(),
            "options": {"blowup_factor": opts.blowup_factor, "fri_number_of_queries": opts.fri_number_of_queries,
                        "coset_offset": opts.coset_offset, "grinding_factor": opts.grinding_factor},
            "proof_bytes": len(proof), "sha256": hashlib.sha256(proof).hexdigest()}


CASES = {"c5": c5, "c5_small": lambda: c5(log_n=16), "c3_20": lambda: c3(20), "c3_24": lambda: c3(24), "c3_26": lambda: c3(26), "c4": c4,
         "c4_small": lambda: c4(log_n=14), "fib70000_80bits": fib_80bits, "fib280000_80bits": lambda: fib_80bits(280000)}


def main():
    names = sys.argv[1:] or ["c5", "c3_24", "fib70000_80bits", "c4", "c3_26"]
    for name in names:
        t0 = time.time()
        res = CASES[name]()
        res["generated_s"] = round(time.time() - t0, 1)
        doc = json.load(open(OUT)) if os.path.exists(OUT) else {"generator": "tests/golden/make_big_configs.py", "cases": {}}
        doc["cases"][name] = res
        json.dump(doc, open(OUT, "w"), indent=1)
        print(name, res, flush=True)


if __name__ == "__main__":
    main()
