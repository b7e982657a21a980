"""DefaultTranscript and the helpers of src/starks/transcript.rs:13-51 (host side, native C++)."""
import ctypes as C

import numpy as np

from . import _native as N


class DefaultTranscript:
    """lambdaworks_crypto::fiat_shamir::default_transcript::DefaultTranscript
    (created at src/starks/prover.rs:91-94): Keccak-256 sponge; `challenge` returns the digest
    byte-reversed and re-seeds the sponge with it."""

    def __init__(self):
        self.handle = C.c_void_p(N.lib().s252_transcript_new())

    def __del__(self):
        if getattr(self, "handle", None):
            N.lib().s252_transcript_free(self.handle)
            self.handle = None

    def append(self, new_data):
        data = bytes(new_data)
        buf = (C.c_uint8 * max(len(data), 1)).from_buffer_copy(data or b"\0")
        N.lib().s252_transcript_append(self.handle, buf, len(data))

    def challenge(self):
        out = (C.c_uint8 * 32)()
        N.lib().s252_transcript_challenge(self.handle, out)
        return bytes(out)


def transcript_to_field(transcript):
    """src/starks/transcript.rs:13-19"""
    out = np.empty(4, dtype=np.uint64)
    N.lib().s252_transcript_to_field(transcript.handle, N.ptr(out))
    return out


def transcript_to_usize(transcript):
    """src/starks/transcript.rs:45-51"""
    return int(N.lib().s252_transcript_to_usize(transcript.handle))


def batch_sample_challenges(size, transcript):
    """src/starks/transcript.rs:71-79"""
    return [transcript_to_field(transcript) for _ in range(size)]
