"""Generates tests/golden/synthetic/commit_vectors.json with the CPU oracle (after the oracle was
pinned on the reference's golden proofs, tests/test_oracle_golden.py).

Inputs are splitmix64-seeded tables (tests/util.py: random_felts(seed, n)); outputs are stored as
Keccak-256 digests of the raw little-endian bytes of the oracle's results, plus the Merkle roots, FRI
roots, last values and nonces themselves.  The GPU box compares the CUDA path against this file.

    python tests/golden/make_synthetic.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import pyoracle as O   # noqa: E402
from util import random_felts      # noqa: E402


def digest(arr):
    return O.keccak256(np.ascontiguousarray(arr).tobytes()).hex()


def commit_case(seed, logn, cols, blowup, offset):
    n = 1 << logn
    trace = random_felts(seed, n * cols).reshape(n, cols, 4)
    r = O.interpolate_and_commit(trace, blowup, offset, threads=8)
    return {"kind": "interpolate_and_commit", "seed": seed, "log_n": logn, "cols": cols, "blowup": blowup, "offset": offset,
            "root": r["root"].hex(), "coeffs_digest": digest(r["coeffs"]), "lde_digest": digest(r["lde"]),
            "nodes_digest": digest(r["nodes"])}


def fri_case(seed, logn, blowup, offset, grinding):
    n = 1 << logn
    p0 = random_felts(seed, n)
    t = O.Transcript()
    t.append(bytes(32))
    last, roots, evals, nodes = O.fri_commit_phase(logn, p0, t, O.fe_from_u64(offset), n * blowup)
    ch = t.challenge()
    return {"kind": "fri_commit_phase", "seed": seed, "log_n": logn, "blowup": blowup, "offset": offset,
            "roots": [x.tobytes().hex() for x in roots], "last_value": O.fe_to_bytes_be(last).hex(),
            "layers_digest": digest(np.concatenate(evals)), "grinding_factor": grinding,
            "nonce": O.generate_nonce_with_grinding(ch, grinding)}


def main():
    cases = [commit_case(0xB200 + 1, 3, 1, 4, 3), commit_case(0xB200 + 2, 10, 34, 4, 3), commit_case(0xB200 + 3, 12, 18, 4, 3),
             commit_case(0xB200 + 4, 11, 33, 8, 3), commit_case(0xB200 + 6, 13, 2, 4, 3), commit_case(0xB200 + 7, 14, 5, 2, 7),
             fri_case(0xB200 + 5, 12, 4, 3, 12), fri_case(0xB200 + 8, 6, 8, 3, 8)]
    out = os.path.join(HERE, "synthetic", "commit_vectors.json")
    json.dump({"generator": "tests/golden/make_synthetic.py", "cases": cases}, open(out, "w"), indent=1)
    print("wrote", out, len(cases), "cases")


if __name__ == "__main__":
    main()
