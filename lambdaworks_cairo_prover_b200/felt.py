"""Stark252 field elements at the API boundary.

`FE` values are numpy uint64[4] in the reference's in-memory layout ("LW"): limbs[0] most
significant, Montgomery form with R = 2^256 (lambdaworks-math U256 Montgomery backend).
Only conversions live here -- the arithmetic of the hot path runs on the GPU.
"""
import numpy as np

MODULUS = 2**251 + 17 * 2**192 + 1
_R = 2**256
_R_INV = pow(_R, -1, MODULUS)
_MASK = 2**64 - 1


def from_int(v):
    """FieldElement::from(v) for a python int (reduced mod p)."""
    m = (int(v) % MODULUS) * _R % MODULUS
    return np.array([(m >> 192) & _MASK, (m >> 128) & _MASK, (m >> 64) & _MASK, m & _MASK], dtype=np.uint64)


def to_int(fe):
    """Canonical representative of an LW element."""
    a = np.asarray(fe, dtype=np.uint64).reshape(4)
    m = (int(a[0]) << 192) | (int(a[1]) << 128) | (int(a[2]) << 64) | int(a[3])
    return m * _R_INV % MODULUS


def from_ints(vs):
    out = np.empty((len(vs), 4), dtype=np.uint64)
    for i, v in enumerate(vs):
        out[i] = from_int(v)
    return out


def to_ints(a):
    return [to_int(x) for x in np.asarray(a, dtype=np.uint64).reshape(-1, 4)]


def to_bytes_be(fe):
    """ByteConversion::to_bytes_be: 32-byte big-endian canonical value."""
    return to_int(fe).to_bytes(32, "big")


def from_bytes_be(b):
    return from_int(int.from_bytes(bytes(b), "big"))


def zero():
    return np.zeros(4, dtype=np.uint64)


def one():
    return from_int(1)


def to_bytes_be_many(a):
    """to_bytes_be for an array of LW elements at once (native): uint64[..., 4] -> uint8[..., 32]."""
    from . import _native as N
    a = np.ascontiguousarray(a, dtype=np.uint64)
    out = np.empty(a.shape[:-1] + (32,), dtype=np.uint8)
    N.lib().s252_fe_to_bytes_be(N.ptr(a), a.size // 4, N.ptr(out))
    return out
