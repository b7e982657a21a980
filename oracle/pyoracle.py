"""ctypes loader for the CPU ORACLE (test infrastructure, NOT the product).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  It builds oracle/libstark252_oracle.so with plain gcc when missing.

Element interchange format ("LW"): numpy uint64 arrays of shape (..., 4), limbs[0] MOST
significant, Montgomery form -- the in-memory FieldElement<Stark252PrimeField> of the reference's
dependency.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libstark252_oracle.so")

P = 2**251 + 17 * 2**192 + 1
R = 2**256


def build(force=False):
    src = os.path.join(_HERE, "stark252_oracle.c")
    srcs = [src, os.path.join(_HERE, "cairo_oracle.inc.c"), os.path.join(_HERE, "stark252_oracle.h")]
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, u64, sz, u8, u32, i32 = C.c_void_p, C.c_uint64, C.c_size_t, C.c_uint8, C.c_uint32, C.c_int
        sig = {
            "o_fe_from_u64": (None, [u64, vp]),
            "o_fe_from_bytes_be": (None, [vp, vp]),
            "o_fe_to_bytes_be": (None, [vp, vp]),
            "o_fe_add": (None, [vp, vp, vp]),
            "o_fe_sub": (None, [vp, vp, vp]),
            "o_fe_mul": (None, [vp, vp, vp]),
            "o_fe_inv": (None, [vp, vp]),
            "o_fe_pow": (None, [vp, u64, vp]),
            "o_primitive_root": (i32, [u32, vp]),
            "o_coset_powers": (i32, [u32, sz, vp, vp]),
            "o_keccak256": (None, [vp, sz, vp]),
            "o_transcript_new": (vp, []),
            "o_transcript_free": (None, [vp]),
            "o_transcript_append": (None, [vp, vp, sz]),
            "o_transcript_challenge": (None, [vp, vp]),
            "o_randomness_to_field": (None, [vp, vp]),
            "o_transcript_to_field": (None, [vp, vp]),
            "o_transcript_to_usize": (u64, [vp]),
            "o_interpolate_fft": (i32, [vp, sz, vp]),
            "o_interpolate_offset_fft": (i32, [vp, sz, vp, vp]),
            "o_evaluate_offset_fft_len": (sz, [vp, sz, sz, sz]),
            "o_evaluate_offset_fft": (i32, [vp, sz, sz, sz, vp, vp]),
            "o_evaluate_polynomial_on_lde_domain": (i32, [vp, sz, sz, sz, vp, vp]),
            "o_merkle_build": (i32, [vp, sz, sz, vp]),
            "o_merkle_path": (i32, [vp, sz, sz, vp]),
            "o_merkle_verify": (i32, [vp, sz, vp, sz, vp, sz]),
            "o_fold_polynomial": (None, [vp, sz, vp, vp]),
            "o_fri_commit_phase": (i32, [sz, vp, sz, vp, vp, sz, vp, vp, vp, vp]),
            "o_grinding_zeros": (u8, [vp, u64]),
            "o_generate_nonce_with_grinding": (i32, [vp, u8, u64, vp]),
            "o_interpolate_and_commit": (i32, [vp, sz, sz, sz, u64, i32, vp, vp, vp, vp]),
            "o_commit_columns": (i32, [vp, sz, sz, vp, vp]),
            "o_poly_evaluate": (None, [vp, sz, vp, vp]),
            "o_deep_composition_poly": (None, [vp, sz, sz, vp, vp, vp, vp, sz, vp, vp, vp, vp, vp, vp, vp]),
            "o_cairo_build_aux_trace": (i32, [vp, sz, sz, vp, vp, sz, vp, vp]),
            "o_cairo_compute_transition": (None, [vp, vp, sz, vp, i32, vp]),
            "o_cairo_constraint_evaluations": (i32, [vp, sz, sz, sz, u64, i32, vp, sz, vp, vp, vp, vp, vp, i32, vp]),
        }
        for name, (res, args) in sig.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _fe_arr(a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    assert a.shape[-1] == 4
    return a


# ---------------------------------------------------------------- conversions (pure python)
def int_to_lw(v):
    """canonical python int -> LW limbs (Montgomery, most-significant limb first)."""
    m = (v % P) * R % P
    return np.array([(m >> (64 * (3 - i))) & (2**64 - 1) for i in range(4)], dtype=np.uint64)


def lw_to_int(a):
    a = np.asarray(a, dtype=np.uint64).reshape(4)
    m = 0
    for i in range(4):
        m = (m << 64) | int(a[i])
    return m * pow(R, -1, P) % P


def ints_to_lw(vs):
    out = np.empty((len(vs), 4), dtype=np.uint64)
    for i, v in enumerate(vs):
        out[i] = int_to_lw(v)
    return out


def lw_to_ints(a):
    a = np.asarray(a, dtype=np.uint64).reshape(-1, 4)
    return [lw_to_int(x) for x in a]


# ---------------------------------------------------------------- field
def fe_from_u64(v):
    out = np.empty(4, dtype=np.uint64)
    lib().o_fe_from_u64(v, _p(out))
    return out


def fe_from_bytes_be(b):
    out = np.empty(4, dtype=np.uint64)
    buf = np.frombuffer(bytes(b), dtype=np.uint8).copy()
    lib().o_fe_from_bytes_be(_p(buf), _p(out))
    return out


def fe_to_bytes_be(a):
    a = _fe_arr(a)
    out = np.empty(32, dtype=np.uint8)
    lib().o_fe_to_bytes_be(_p(a), _p(out))
    return out.tobytes()


def _binop(name, a, b):
    a, b = _fe_arr(a), _fe_arr(b)
    out = np.empty(4, dtype=np.uint64)
    getattr(lib(), name)(_p(a), _p(b), _p(out))
    return out


def fe_add(a, b):
    return _binop("o_fe_add", a, b)


def fe_sub(a, b):
    return _binop("o_fe_sub", a, b)


def fe_mul(a, b):
    return _binop("o_fe_mul", a, b)


def fe_inv(a):
    a = _fe_arr(a)
    out = np.empty(4, dtype=np.uint64)
    lib().o_fe_inv(_p(a), _p(out))
    return out


def fe_pow(a, e):
    a = _fe_arr(a)
    out = np.empty(4, dtype=np.uint64)
    lib().o_fe_pow(_p(a), e, _p(out))
    return out


def primitive_root(order):
    out = np.empty(4, dtype=np.uint64)
    assert lib().o_primitive_root(order, _p(out)) == 0
    return out


def coset_powers(order, count, offset):
    offset = _fe_arr(offset)
    out = np.empty((count, 4), dtype=np.uint64)
    assert lib().o_coset_powers(order, count, _p(offset), _p(out)) == 0
    return out


# ---------------------------------------------------------------- keccak / transcript
def keccak256(data):
    buf = np.frombuffer(bytes(data), dtype=np.uint8).copy() if len(data) else np.zeros(1, dtype=np.uint8)
    out = np.empty(32, dtype=np.uint8)
    lib().o_keccak256(_p(buf), len(data), _p(out))
    return out.tobytes()


def randomness_to_field(b):
    buf = np.frombuffer(bytes(b), dtype=np.uint8).copy()
    out = np.empty(4, dtype=np.uint64)
    lib().o_randomness_to_field(_p(buf), _p(out))
    return out


class Transcript:
    """DefaultTranscript of the reference's dependency."""

    def __init__(self):
        self._t = lib().o_transcript_new()

    def __del__(self):
        if getattr(self, "_t", None):
            lib().o_transcript_free(self._t)
            self._t = None

    @property
    def handle(self):
        return self._t

    def append(self, data):
        data = bytes(data)
        buf = np.frombuffer(data, dtype=np.uint8).copy() if data else np.zeros(1, dtype=np.uint8)
        lib().o_transcript_append(self._t, _p(buf), len(data))

    def challenge(self):
        out = np.empty(32, dtype=np.uint8)
        lib().o_transcript_challenge(self._t, _p(out))
        return out.tobytes()

    def to_field(self):
        out = np.empty(4, dtype=np.uint64)
        lib().o_transcript_to_field(self._t, _p(out))
        return out

    def to_usize(self):
        return int(lib().o_transcript_to_usize(self._t))


# ---------------------------------------------------------------- FFTPoly
def interpolate_fft(evals):
    evals = _fe_arr(evals)
    out = np.empty_like(evals)
    rc = lib().o_interpolate_fft(_p(evals), evals.shape[0], _p(out))
    if rc:
        raise ValueError("FFTError: input length is not a power of two")
    return out


def interpolate_offset_fft(evals, offset):
    evals, offset = _fe_arr(evals), _fe_arr(offset)
    out = np.empty_like(evals)
    rc = lib().o_interpolate_offset_fft(_p(evals), evals.shape[0], _p(offset), _p(out))
    if rc:
        raise ValueError("FFTError: input length is not a power of two")
    return out


def evaluate_offset_fft(coeffs, blowup, domain_size, offset):
    coeffs, offset = _fe_arr(coeffs).reshape(-1, 4), _fe_arr(offset)
    ds = domain_size or 0
    n = lib().o_evaluate_offset_fft_len(_p(coeffs), coeffs.shape[0], blowup, ds)
    out = np.empty((n, 4), dtype=np.uint64)
    rc = lib().o_evaluate_offset_fft(_p(coeffs), coeffs.shape[0], blowup, ds, _p(offset), _p(out))
    if rc:
        raise ValueError("FFTError")
    return out


def evaluate_polynomial_on_lde_domain(coeffs, blowup, domain_size, offset):
    coeffs, offset = _fe_arr(coeffs).reshape(-1, 4), _fe_arr(offset)
    out = np.empty((domain_size * blowup, 4), dtype=np.uint64)
    rc = lib().o_evaluate_polynomial_on_lde_domain(_p(coeffs), coeffs.shape[0], blowup, domain_size,
                                                   _p(offset), _p(out))
    if rc:
        raise ValueError("FFTError")
    return out


# ---------------------------------------------------------------- Merkle
def merkle_build(rows):
    """rows: (n_leaves, n_cols, 4) LW -> nodes (2n-1, 32) uint8, heap layout, root at 0."""
    rows = _fe_arr(rows)
    if rows.ndim == 2:
        rows = rows.reshape(rows.shape[0], 1, 4)
    n, c = rows.shape[0], rows.shape[1]
    nodes = np.empty((2 * n - 1, 32), dtype=np.uint8)
    rc = lib().o_merkle_build(_p(rows), n, c, _p(nodes))
    if rc:
        raise ValueError("merkle_build: leaf count must be a power of two")
    return nodes


def merkle_path(nodes, pos):
    n = (nodes.shape[0] + 1) // 2
    k = n.bit_length() - 1
    path = np.empty((k, 32), dtype=np.uint8)
    rc = lib().o_merkle_path(_p(nodes), n, pos, _p(path))
    if rc:
        return None
    return path


def merkle_verify(root, index, value, path):
    value = _fe_arr(value).reshape(-1, 4)
    root = np.frombuffer(bytes(root), dtype=np.uint8).copy()
    if isinstance(path, (list, tuple)):
        path = np.frombuffer(b"".join(bytes(x) for x in path), dtype=np.uint8).copy()
    path = np.ascontiguousarray(path, dtype=np.uint8).reshape(-1, 32)
    return bool(lib().o_merkle_verify(_p(root), index, _p(value), value.shape[0], _p(path), path.shape[0]))


def commit_columns(cols):
    """cols: (n_cols, n_rows, 4) column-major -> (nodes, root)."""
    cols = _fe_arr(cols)
    c, n = cols.shape[0], cols.shape[1]
    nodes = np.empty((2 * n - 1, 32), dtype=np.uint8)
    root = np.empty(32, dtype=np.uint8)
    assert lib().o_commit_columns(_p(cols), n, c, _p(nodes), _p(root)) == 0
    return nodes, root.tobytes()


# ---------------------------------------------------------------- FRI / grinding
def fold_polynomial(coeffs, beta):
    coeffs, beta = _fe_arr(coeffs).reshape(-1, 4), _fe_arr(beta)
    n = coeffs.shape[0]
    out = np.empty(((n + 1) // 2, 4), dtype=np.uint64)
    lib().o_fold_polynomial(_p(coeffs), n, _p(beta), _p(out))
    return out


def fri_commit_phase(number_layers, p0, transcript, coset_offset, domain_size, keep=True):
    """Returns (last_value, roots[(L,32)], layer_evals[list], layer_nodes[list])."""
    p0, coset_offset = _fe_arr(p0).reshape(-1, 4), _fe_arr(coset_offset)
    evals = [np.empty((domain_size >> k, 4), dtype=np.uint64) for k in range(number_layers)] if keep else None
    nodes = [np.empty((2 * (domain_size >> k) - 1, 32), dtype=np.uint8) for k in range(number_layers)] if keep else None
    ev_ptrs = (C.c_void_p * max(number_layers, 1))(*[e.ctypes.data for e in evals]) if keep else None
    nd_ptrs = (C.c_void_p * max(number_layers, 1))(*[e.ctypes.data for e in nodes]) if keep else None
    roots = np.empty((number_layers, 32), dtype=np.uint8)
    last = np.empty(4, dtype=np.uint64)
    rc = lib().o_fri_commit_phase(number_layers, _p(p0), p0.shape[0], transcript.handle, _p(coset_offset),
                                  domain_size, _p(last), ev_ptrs, nd_ptrs, _p(roots))
    if rc:
        raise ValueError("fri_commit_phase failed rc=%d" % rc)
    return last, roots, evals, nodes


def grinding_zeros(challenge, nonce):
    ch = np.frombuffer(bytes(challenge), dtype=np.uint8).copy()
    return int(lib().o_grinding_zeros(_p(ch), nonce))


def generate_nonce_with_grinding(challenge, grinding_factor, limit=2**64 - 1):
    ch = np.frombuffer(bytes(challenge), dtype=np.uint8).copy()
    nonce = C.c_uint64(0)
    ok = lib().o_generate_nonce_with_grinding(_p(ch), grinding_factor, limit, C.byref(nonce))
    return int(nonce.value) if ok else None


# ---------------------------------------------------------------- interpolate_and_commit
def interpolate_and_commit(trace, blowup, coset_offset, threads=1, want_lde=True, want_nodes=True):
    """trace: (n_rows, n_cols, 4) row-major.  Returns dict(coeffs, lde, nodes, root)."""
    trace = _fe_arr(trace)
    n, c = trace.shape[0], trace.shape[1]
    m = n * blowup
    coeffs = np.empty((c, n, 4), dtype=np.uint64)
    lde = np.empty((c, m, 4), dtype=np.uint64) if want_lde else None
    nodes = np.empty((2 * m - 1, 32), dtype=np.uint8) if want_nodes else None
    root = np.empty(32, dtype=np.uint8)
    rc = lib().o_interpolate_and_commit(_p(trace), n, c, blowup, coset_offset, threads, _p(coeffs), _p(lde),
                                        _p(nodes), _p(root))
    if rc:
        raise ValueError("interpolate_and_commit failed rc=%d" % rc)
    return {"coeffs": coeffs, "lde": lde, "nodes": nodes, "root": root.tobytes()}


# ---------------------------------------------------------------- round 3 / round 4 helpers
def poly_evaluate(coeffs, x):
    coeffs, x = _fe_arr(coeffs).reshape(-1, 4), _fe_arr(x)
    out = np.empty(4, dtype=np.uint64)
    lib().o_poly_evaluate(_p(coeffs), coeffs.shape[0], _p(x), _p(out))
    return out


def deep_composition_poly(trace_polys, h1, h2, z, offsets, ood, h1_z2, h2_z2, gamma, gamma_p, gammas):
    """trace_polys (c, n, 4); h1, h2 (n, 4); ood (K, c, 4); gammas (c, K, 4) -> p0 coefficients (n, 4)."""
    trace_polys = _fe_arr(trace_polys)
    c, n = trace_polys.shape[0], trace_polys.shape[1]
    offs = np.ascontiguousarray(offsets, dtype=np.uint64)
    out = np.empty((n, 4), dtype=np.uint64)
    lib().o_deep_composition_poly(_p(trace_polys), c, n, _p(_fe_arr(h1)), _p(_fe_arr(h2)), _p(_fe_arr(z)), _p(offs), len(offs),
                                  _p(_fe_arr(ood)), _p(_fe_arr(h1_z2)), _p(_fe_arr(h2_z2)), _p(_fe_arr(gamma)),
                                  _p(_fe_arr(gamma_p)), _p(_fe_arr(gammas)), _p(out))
    return out


# ---------------------------------------------------------------- Cairo AIR (cairo_oracle.inc.c)
def cairo_build_aux_trace(main, pub_addrs, pub_vals, rap):
    """main (n, c, 4) row-major; public memory in address order; rap (3, 4) -> aux (n, 18, 4)."""
    main = _fe_arr(main)
    n, c = main.shape[0], main.shape[1]
    addrs = np.ascontiguousarray(pub_addrs, dtype=np.uint64)
    vals = _fe_arr(pub_vals).reshape(-1, 4)
    out = np.empty((n, 18, 4), dtype=np.uint64)
    rc = lib().o_cairo_build_aux_trace(_p(main), n, c, _p(addrs), _p(vals), len(addrs), _p(_fe_arr(rap)), _p(out))
    if rc:
        raise ValueError("cairo_build_aux_trace failed rc=%d" % rc)
    return out


def cairo_compute_transition(cur, nxt, rap, has_rc=False):
    cur, nxt = _fe_arr(cur).reshape(-1, 4), _fe_arr(nxt).reshape(-1, 4)
    out = np.empty((50 if has_rc else 49, 4), dtype=np.uint64)
    lib().o_cairo_compute_transition(_p(cur), _p(nxt), cur.shape[0], _p(_fe_arr(rap)), int(has_rc), _p(out))
    return out


def cairo_constraint_evaluations(lde, n, blowup, coset_offset, rap, boundary, boundary_coeffs, transition_coeffs, has_rc=False,
                                 threads=1):
    """lde (c, m, 4) column-major; boundary = [(col, step, value LW)]; coeffs (k, 2, 4) -> (m, 4)."""
    lde = _fe_arr(lde)
    c, m = lde.shape[0], lde.shape[1]
    assert m == n * blowup
    bcols = np.array([b[0] for b in boundary], dtype=np.uint64)
    bsteps = np.array([b[1] for b in boundary], dtype=np.uint64)
    bvals = _fe_arr(np.stack([b[2] for b in boundary]))
    out = np.empty((m, 4), dtype=np.uint64)
    rc = lib().o_cairo_constraint_evaluations(_p(lde), c, n, blowup, coset_offset, int(has_rc), _p(_fe_arr(rap)), len(boundary),
                                              _p(bcols), _p(bsteps), _p(bvals), _p(_fe_arr(boundary_coeffs)),
                                              _p(_fe_arr(transition_coeffs)), threads, _p(out))
    if rc:
        raise ValueError("cairo_constraint_evaluations failed rc=%d" % rc)
    return out
