"""FRI commit and query phases (src/starks/fri/mod.rs) on the GPU."""
import ctypes as C

import numpy as np

from . import _native as N
from .merkle import Proof


class FriLayer:
    """View of one layer of a device-resident FRI commitment
    (src/starks/fri/fri_commitment.rs:14-24: evaluation, merkle_tree, coset_offset, domain_size)."""

    def __init__(self, owner, index, domain_size, root):
        self._owner, self.index, self.domain_size, self.root = owner, index, domain_size, root

    @property
    def evaluation(self):
        out = np.empty((self.domain_size, 4), dtype=np.uint64)
        self._owner.ctx.check(N.lib().s252_fri_read_layer(self._owner.handle, self.index, 0, self.domain_size, N.ptr(out)))
        return out

    def nodes(self):
        out = np.empty((2 * self.domain_size - 1, 32), dtype=np.uint8)
        self._owner.ctx.check(N.lib().s252_fri_read_nodes(self._owner.handle, self.index, 0, out.shape[0], N.ptr(out)))
        return out


class FriLayers:
    def __init__(self, ctx, handle, domain_size, roots):
        self.ctx, self.handle, self.domain_size = ctx, handle, domain_size
        self.layers = [FriLayer(self, k, domain_size >> k, roots[k].tobytes()) for k in range(roots.shape[0])]
        ctx.adopt(self)

    def __len__(self):
        return len(self.layers)

    def __iter__(self):
        return iter(self.layers)

    def __getitem__(self, k):
        return self.layers[k]

    def free(self):
        if getattr(self, "handle", None):
            N.lib().s252_fri_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def fri_commit_phase(number_layers, p_0, transcript, coset_offset, domain_size, ctx=None):
    """src/starks/fri/mod.rs:20-72 -> (last_value, fri_layer_list)

    p_0: Polynomial (or coefficient array); transcript: DefaultTranscript, advanced exactly as the
    reference does (append root_k, sample zeta_k, ..., append last value)."""
    ctx = ctx or N.default_context()
    coeffs = getattr(p_0, "coefficients", p_0)
    coeffs = N.fe_array(np.asarray(coeffs, dtype=np.uint64).reshape(-1, 4))
    off = N.fe_array(coset_offset)
    h = C.c_void_p()
    last = np.empty(4, dtype=np.uint64)
    roots = np.empty((number_layers, 32), dtype=np.uint8)
    ctx.check(N.lib().s252_fri_commit_phase(ctx.handle, number_layers, N.ptr(coeffs) if coeffs.shape[0] else None,
                                            coeffs.shape[0], transcript.handle, N.ptr(off), domain_size, N.HOST,
                                            C.byref(h), N.ptr(last), N.ptr(roots) if number_layers else None))
    return last, FriLayers(ctx, h, domain_size, roots)


class FriDecommitment:
    """src/starks/fri/fri_decommit.rs:11-17"""

    def __init__(self, layers_auth_paths_sym, layers_evaluations_sym, layers_auth_paths, layers_evaluations):
        self.layers_auth_paths_sym = layers_auth_paths_sym
        self.layers_evaluations_sym = layers_evaluations_sym
        self.layers_auth_paths = layers_auth_paths
        self.layers_evaluations = layers_evaluations


def fri_query_phase(number_of_queries, domain_size, fri_layers, transcript):
    """src/starks/fri/mod.rs:74-127 -> (query_list, iotas)"""
    from .transcript import transcript_to_usize
    if len(fri_layers) == 0:
        return [], []
    iotas = [transcript_to_usize(transcript) % domain_size for _ in range(number_of_queries)]
    return fri_open(fri_layers, iotas), iotas


def fri_open(fri_layers, iotas):
    ctx = fri_layers.ctx
    L, Q = len(fri_layers), len(iotas)
    stride = max(domain_bits(fri_layers.domain_size), 1)
    idx = np.array(iotas, dtype=np.uint64)
    evals = np.empty((Q, L, 4), dtype=np.uint64)
    evals_sym = np.empty((Q, L, 4), dtype=np.uint64)
    paths = np.zeros((Q, L, stride, 32), dtype=np.uint8)
    paths_sym = np.zeros((Q, L, stride, 32), dtype=np.uint8)
    ctx.check(N.lib().s252_fri_query(fri_layers.handle, N.ptr(idx), Q, N.ptr(evals), N.ptr(evals_sym), N.ptr(paths),
                                     N.ptr(paths_sym), stride))
    out = []
    for q in range(Q):
        depth = [domain_bits(fri_layers.domain_size >> k) for k in range(L)]
        out.append(FriDecommitment(
            [Proof([paths_sym[q, k, d].tobytes() for d in range(depth[k])]) for k in range(L)],
            [evals_sym[q, k].copy() for k in range(L)],
            [Proof([paths[q, k, d].tobytes() for d in range(depth[k])]) for k in range(L)],
            [evals[q, k].copy() for k in range(L)]))
    return out


def domain_bits(n):
    return n.bit_length() - 1
