"""ProofOptions and its presets (src/starks/proof/options.rs:21-74, 144-151)."""
from dataclasses import dataclass


@dataclass
class ProofOptions:
    blowup_factor: int
    fri_number_of_queries: int
    coset_offset: int
    grinding_factor: int

    @staticmethod
    def default_test_options():
        return ProofOptions(4, 3, 3, 1)

    @staticmethod
    def new_secure(security_level, coset_offset):
        queries = {"Conjecturable80Bits": 31, "Conjecturable100Bits": 41, "Conjecturable128Bits": 55,
                   "Provable80Bits": 80, "Provable100Bits": 104, "Provable128Bits": 140}[security_level]
        return ProofOptions(4, queries, coset_offset, 20)
