"""Wire format of the reference's proof files (ORACLE side: test infrastructure, NOT the product).

Restates, for parsing and re-serialising:
  * the proof-file container            src/main.rs:98-102   u64_be(len) || StarkProof || PublicInputs
  * StarkProof::{serialize,deserialize} src/starks/proof/stark.rs:161-218, 220-420
  * DeepPolynomialOpenings              src/starks/proof/stark.rs:54-80
  * Frame                               src/starks/frame.rs:86-105
  * FriDecommitment                     src/starks/fri/fri_decommit.rs:19-45
  * Proof<Commitment> (Merkle path)     src/starks/utils.rs:6-13
All integers are big-endian u64; field elements are 32-byte big-endian canonical values.
"""
from dataclasses import dataclass, field
from typing import List


class Reader:
    def __init__(self, b):
        self.b = memoryview(bytes(b))
        self.o = 0

    def u64(self):
        if self.o + 8 > len(self.b):
            raise ValueError("InvalidAmountOfBytes")
        v = int.from_bytes(self.b[self.o:self.o + 8], "big")
        self.o += 8
        return v

    def take(self, n):
        if self.o + n > len(self.b):
            raise ValueError("InvalidAmountOfBytes")
        v = bytes(self.b[self.o:self.o + n])
        self.o += n
        return v

    def felt(self, felt_len=32):
        return int.from_bytes(self.take(felt_len), "big")

    def rest(self):
        return bytes(self.b[self.o:])


def _u64(v):
    return int(v).to_bytes(8, "big")


def _felt(v):
    return int(v).to_bytes(32, "big")


def read_merkle_path(r):
    n = r.u64()
    return [r.take(32) for _ in range(n)]


def write_merkle_path(path):
    return _u64(len(path)) + b"".join(path)


@dataclass
class Frame:
    data: List[int]
    row_width: int

    def num_rows(self):
        return len(self.data) // self.row_width if self.row_width else 0

    def row(self, i):
        return self.data[i * self.row_width:(i + 1) * self.row_width]

    @staticmethod
    def parse(b):
        r = Reader(b)
        n = r.u64()
        felt_len = r.u64()
        data = [r.felt(felt_len) for _ in range(n)]
        row_width = r.u64()
        return Frame(data, row_width)

    def serialize(self):
        out = _u64(len(self.data)) + _u64(32 if self.data else 0)
        out += b"".join(_felt(v) for v in self.data)
        return out + _u64(self.row_width)


@dataclass
class FriDecommitment:
    layers_auth_paths_sym: List[List[bytes]]
    layers_evaluations_sym: List[int]
    layers_evaluations: List[int]
    layers_auth_paths: List[List[bytes]]

    @staticmethod
    def parse(b):
        r = Reader(b)
        n = r.u64()
        paths_sym = [read_merkle_path(r) for _ in range(n)]
        felt_len = r.u64()
        n = r.u64()
        evs_sym = [r.felt(felt_len) for _ in range(n)]
        n = r.u64()
        evs = [r.felt(felt_len) for _ in range(n)]
        n = r.u64()
        paths = [read_merkle_path(r) for _ in range(n)]
        return FriDecommitment(paths_sym, evs_sym, evs, paths)

    def serialize(self):
        out = _u64(len(self.layers_auth_paths_sym))
        out += b"".join(write_merkle_path(p) for p in self.layers_auth_paths_sym)
        out += _u64(32)
        out += _u64(len(self.layers_evaluations_sym)) + b"".join(_felt(v) for v in self.layers_evaluations_sym)
        out += _u64(len(self.layers_evaluations)) + b"".join(_felt(v) for v in self.layers_evaluations)
        out += _u64(len(self.layers_auth_paths))
        out += b"".join(write_merkle_path(p) for p in self.layers_auth_paths)
        return out


@dataclass
class DeepPolynomialOpenings:
    lde_composition_poly_proof: List[bytes]
    lde_composition_poly_even_evaluation: int
    lde_composition_poly_odd_evaluation: int
    lde_trace_merkle_proofs: List[List[bytes]]
    lde_trace_evaluations: List[int]

    @staticmethod
    def parse(b):
        r = Reader(b)
        comp_path = read_merkle_path(r)
        felt_len = r.u64()
        even = r.felt(felt_len)
        odd = r.felt(felt_len)
        n = r.u64()
        proofs = [read_merkle_path(r) for _ in range(n)]
        n = r.u64()
        evs = [r.felt(felt_len) for _ in range(n)]
        return DeepPolynomialOpenings(comp_path, even, odd, proofs, evs)

    def serialize(self):
        out = write_merkle_path(self.lde_composition_poly_proof)
        out += _u64(32) + _felt(self.lde_composition_poly_even_evaluation)
        out += _felt(self.lde_composition_poly_odd_evaluation)
        out += _u64(len(self.lde_trace_merkle_proofs))
        out += b"".join(write_merkle_path(p) for p in self.lde_trace_merkle_proofs)
        out += _u64(len(self.lde_trace_evaluations)) + b"".join(_felt(v) for v in self.lde_trace_evaluations)
        return out


@dataclass
class StarkProof:
    trace_length: int
    lde_trace_merkle_roots: List[bytes]
    trace_ood_frame_evaluations: Frame
    composition_poly_root: bytes
    composition_poly_even_ood_evaluation: int
    composition_poly_odd_ood_evaluation: int
    fri_layers_merkle_roots: List[bytes]
    fri_last_value: int
    query_list: List[FriDecommitment]
    deep_poly_openings: List[DeepPolynomialOpenings]
    nonce: int
    trailing: bytes = field(default=b"", repr=False)

    @staticmethod
    def parse(b):
        r = Reader(b)
        trace_length = r.u64()
        n = r.u64()
        roots = [r.take(32) for _ in range(n)]
        n = r.u64()
        frame = Frame.parse(r.take(n))
        comp_root = r.take(32)
        felt_len = r.u64()
        even = r.felt(felt_len)
        odd = r.felt(felt_len)
        n = r.u64()
        fri_roots = [r.take(32) for _ in range(n)]
        last = r.felt(felt_len)
        n = r.u64()
        queries = []
        for _ in range(n):
            ln = r.u64()
            queries.append(FriDecommitment.parse(r.take(ln)))
        n = r.u64()
        openings = []
        for _ in range(n):
            ln = r.u64()
            openings.append(DeepPolynomialOpenings.parse(r.take(ln)))
        nonce = r.u64()
        return StarkProof(trace_length, roots, frame, comp_root, even, odd, fri_roots, last, queries,
                          openings, nonce, r.rest())

    def serialize(self):
        out = _u64(self.trace_length)
        out += _u64(len(self.lde_trace_merkle_roots)) + b"".join(self.lde_trace_merkle_roots)
        fb = self.trace_ood_frame_evaluations.serialize()
        out += _u64(len(fb)) + fb
        out += self.composition_poly_root
        out += _u64(32) + _felt(self.composition_poly_even_ood_evaluation)
        out += _felt(self.composition_poly_odd_ood_evaluation)
        out += _u64(len(self.fri_layers_merkle_roots)) + b"".join(self.fri_layers_merkle_roots)
        out += _felt(self.fri_last_value)
        out += _u64(len(self.query_list))
        for q in self.query_list:
            qb = q.serialize()
            out += _u64(len(qb)) + qb
        out += _u64(len(self.deep_poly_openings))
        for o in self.deep_poly_openings:
            ob = o.serialize()
            out += _u64(len(ob)) + ob
        out += _u64(self.nonce)
        return out


def read_proof_file(path):
    """src/main.rs:98-102: u64_be(len) || StarkProof::serialize() || PublicInputs::serialize()."""
    raw = open(path, "rb").read()
    n = int.from_bytes(raw[:8], "big")
    proof_bytes = raw[8:8 + n]
    return StarkProof.parse(proof_bytes), proof_bytes, raw[8 + n:]
