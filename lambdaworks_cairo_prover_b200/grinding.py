"""Proof-of-work nonce search (src/starks/grinding.rs) on the GPU."""
import ctypes as C

import numpy as np

from . import _native as N


def generate_nonce_with_grinding(transcript_challenge, grinding_factor, ctx=None, limit=0):
    """src/starks/grinding.rs:40-48 -> Option<u64>: the smallest nonce whose
    Keccak256(challenge || nonce_le) head has >= grinding_factor trailing zero bits, or None."""
    ctx = ctx or N.default_context()
    ch = np.frombuffer(bytes(transcript_challenge), dtype=np.uint8).copy()
    if ch.shape[0] != 32:
        raise ValueError("transcript challenge is 32 bytes")
    nonce = C.c_uint64(0)
    rc = N.lib().s252_generate_nonce_with_grinding(ctx.handle, N.ptr(ch), grinding_factor, limit, C.byref(nonce))
    if rc == N.ERR_NOT_FOUND:
        return None
    ctx.check(rc)
    return int(nonce.value)
