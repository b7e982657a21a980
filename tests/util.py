"""Seeded synthetic inputs shared by the tests (splitmix64-seeded limbs
reduced mod p)."""
import numpy as np

from lambdaworks_cairo_prover_b200 import felt

P = felt.MODULUS
_M64 = 2**64 - 1


def splitmix64_stream(seed, n):
    """n 64-bit outputs of splitmix64 as a numpy uint64 array (vectorised)."""
    idx = np.arange(1, n + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def random_felts(seed, n):
    """n field elements in LW layout (uint64[n,4]): random Montgomery residues below p.
    The top limb is masked to 59 bits so every value is < 2^251 < p (no rejection needed)."""
    raw = splitmix64_stream(seed, 4 * n).reshape(n, 4).copy()
    raw[:, 0] &= np.uint64((1 << 59) - 1)
    return raw


def edge_felts():
    vals = [0, 1, 2, P - 1, P - 2, 2**251, 2**192, 2**192 - 1, (P - 1) // 2, 2**64, 2**128 + 12345]
    return felt.from_ints(vals)
