"""ctypes binding of libstark252_b200.so (the C ABI in include/stark252_b200.h).

There is no CPU fallback: creating a Context without a CUDA device raises.
"""
import ctypes as C
import os
import weakref

import numpy as np

from . import build as _build

OK, ERR_INVALID, ERR_CUDA, ERR_NOT_FOUND, ERR_RANGE = 0, -1, -2, -3, -4
HOST, DEVICE = 0, 1

_lib = None

# every symbol include/stark252_b200.h declares: name -> (restype, argtypes)
_vp, _u64, _sz, _i, _u8 = C.c_void_p, C.c_uint64, C.c_size_t, C.c_int, C.c_uint8
SIGNATURES = {
    "s252_ctx_create": (_i, [_i, C.POINTER(_vp)]),
    "s252_ctx_destroy": (None, [_vp]),
    "s252_last_error": (C.c_char_p, [_vp]),
    "s252_ctx_synchronize": (_i, [_vp]),
    "s252_ctx_stream": (_vp, [_vp]),
    "s252_ctx_launch_count": (_u64, [_vp]),
    "s252_ctx_trim": (_i, [_vp]),
    "s252_ctx_profile": (_i, [_vp, _i]),
    "s252_ctx_profile_read": (_i, [_vp, C.c_char_p, _sz]),
    "s252_device_alloc": (_i, [_vp, _sz, C.POINTER(_vp)]),
    "s252_device_free": (_i, [_vp, _vp]),
    "s252_copy_to_device": (_i, [_vp, _vp, _vp, _sz]),
    "s252_copy_to_host": (_i, [_vp, _vp, _vp, _sz]),
    "s252_host_register": (_i, [_vp, _sz]),
    "s252_host_unregister": (_i, [_vp]),
    "s252_copy_2d_to_device": (_i, [_vp, _vp, _sz, _vp, _sz, _sz, _sz]),
    "s252_copy_to_device_async": (_i, [_vp, _vp, _vp, _sz]),
    "s252_copy_stream_wait": (_i, [_vp]),
    "s252_interpolate_fft": (_i, [_vp, _vp, _sz, _vp, _i]),
    "s252_interpolate_offset_fft": (_i, [_vp, _vp, _sz, _vp, _vp, _i]),
    "s252_evaluate_offset_fft_len": (_sz, [_sz, _sz, _sz]),
    "s252_evaluate_offset_fft": (_i, [_vp, _vp, _sz, _sz, _sz, _vp, _vp, _sz, _i]),
    "s252_evaluate_polynomial_on_lde_domain": (_i, [_vp, _vp, _sz, _sz, _sz, _vp, _vp, _i]),
    "s252_ntt_shared": (_i, [_vp, C.c_uint, _i, _sz, _u64, _i, C.c_uint, C.c_uint, _vp, _vp, _vp, C.POINTER(C.c_uint)]),
    "s252_convert_elements": (_i, [_vp, _vp, _vp, _sz, _i]),
    "s252_interpolate_and_commit": (_i, [_vp, _vp, _sz, _sz, _sz, _u64, _i, C.POINTER(_vp), _vp]),
    "s252_interpolate_and_lde": (_i, [_vp, _vp, _sz, _sz, _sz, _u64, _i, C.POINTER(_vp)]),
    "s252_commit_device_columns": (_i, [_vp, _vp, _sz, _sz, _sz, C.POINTER(_vp), _vp]),
    "s252_commit_device_columns_inplace": (_i, [_vp, _vp, _sz, _sz, _sz, C.POINTER(_vp), _vp]),
    "s252_lde_and_commit": (_i, [_vp, _vp, _sz, _sz, _sz, _sz, _u64, _i, C.POINTER(_vp), _vp]),
    "s252_merkle_build": (_i, [_vp, _vp, _sz, _sz, _i, C.POINTER(_vp), _vp]),
    "s252_commit_destroy": (None, [_vp]),
    "s252_commit_n_cols": (_sz, [_vp]),
    "s252_commit_n_rows": (_sz, [_vp]),
    "s252_commit_n_coeffs": (_sz, [_vp]),
    "s252_commit_root": (_i, [_vp, _vp]),
    "s252_commit_read_lde": (_i, [_vp, _sz, _sz, _sz, _vp]),
    "s252_commit_read_coeffs": (_i, [_vp, _sz, _vp]),
    "s252_commit_read_nodes": (_i, [_vp, _sz, _sz, _vp]),
    "s252_cairo_prove_sharded": (_i, [_vp, _vp, _vp, _sz, _sz, _u64, C.c_uint8, _sz, C.POINTER(_vp), C.POINTER(_sz)]),
    "s252_comm_unique_id": (_i, [_vp]),
    "s252_comm_create": (_i, [_vp, _vp, _i, _i, C.POINTER(_vp)]),
    "s252_comm_destroy": (None, [_vp]),
    "s252_comm_rank": (_i, [_vp]),
    "s252_comm_world": (_i, [_vp]),
    "s252_interpolate_and_commit_sharded": (_i, [_vp, _vp, _vp, _vp, _sz, _sz, _sz, _sz, _u64, _i, C.POINTER(_vp), _vp]),
    "s252_sharded_commit_destroy": (None, [_vp]),
    "s252_sharded_commit_n_rows": (_sz, [_vp]),
    "s252_sharded_commit_n_cols": (_sz, [_vp]),
    "s252_sharded_commit_n_local": (_sz, [_vp]),
    "s252_sharded_commit_local": (_vp, [_vp, _sz]),
    "s252_sharded_commit_block": (_vp, [_vp]),
    "s252_sharded_commit_open": (_i, [_vp, _vp, _sz, _vp, _vp]),
    "s252_commit_open": (_i, [_vp, _vp, _sz, _vp, _vp]),
    "s252_commit_device_lde": (_vp, [_vp]),
    "s252_commit_device_coeffs": (_vp, [_vp]),
    "s252_commit_device_nodes": (_vp, [_vp]),
    "s252_fri_commit_phase": (_i, [_vp, _sz, _vp, _sz, _vp, _vp, _sz, _i, C.POINTER(_vp), _vp, _vp]),
    "s252_fri_layer0": (_i, [_vp, _vp, _sz, _vp, _sz, _i, C.POINTER(_vp), _vp]),
    "s252_fri_fold_commit": (_i, [_vp, _vp, _vp]),
    "s252_fri_fold_last": (_i, [_vp, _vp, _vp]),
    "s252_commit_evaluate_at": (_i, [_vp, _vp, _sz, _vp, _sz, _sz]),
    "s252_fri_commit_phase_deep": (_i, [_vp, _sz, _vp, _sz, _vp, _vp, _vp, _sz, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _u64,
                                        C.POINTER(_vp), _vp, _vp]),
    "s252_fri_commit_phase_evals": (_i, [_vp, _sz, _vp, _sz, _vp, _u64, C.POINTER(_vp), _vp, _vp]),
    "s252_fri_fold_rows": (_i, [_vp, _vp, _vp, _sz, _sz, _sz, _sz, _sz, _vp, _u64, _vp]),
    "s252_fri_commit_phase_from_layer": (_i, [_vp, _sz, _vp, _sz, _vp, _u64, _sz, C.POINTER(_vp), _vp, _vp]),
    "s252_fri_destroy": (None, [_vp]),
    "s252_fri_n_layers": (_sz, [_vp]),
    "s252_fri_read_layer": (_i, [_vp, _sz, _sz, _sz, _vp]),
    "s252_fri_read_nodes": (_i, [_vp, _sz, _sz, _sz, _vp]),
    "s252_fri_query": (_i, [_vp, _vp, _sz, _vp, _vp, _vp, _vp, _sz]),
    "s252_generate_nonce_with_grinding": (_i, [_vp, _vp, _u8, _u64, C.POINTER(_u64)]),
    "s252_grind_round": (_i, [_vp, _vp, _u8, _u64, _u64, C.c_uint, C.c_uint, C.c_uint, C.POINTER(_u64)]),
    "s252_fe_to_bytes_be": (None, [_vp, _sz, _vp]),
    "s252_keccak256": (None, [_vp, _sz, _vp]),
    "s252_transcript_new": (_vp, []),
    "s252_transcript_free": (None, [_vp]),
    "s252_transcript_append": (None, [_vp, _vp, _sz]),
    "s252_transcript_challenge": (None, [_vp, _vp]),
    "s252_transcript_to_field": (None, [_vp, _vp]),
    "s252_transcript_to_usize": (_u64, [_vp]),
    "s252_microbench_int_pipes": (_i, [_vp, _vp]),
    "s252_microbench_fe_mul": (_i, [_vp, C.POINTER(C.c_double)]),
    "s252_microbench_keccak": (_i, [_vp, C.POINTER(C.c_double)]),
    "s252_fe_binop": (_i, [_vp, _i, _vp, _vp, _vp, _sz, _i]),
    "s252_keccak256_batch": (_i, [_vp, _vp, _sz, _sz, _vp]),
    # include/stark252_cairo.h
    "s252_cairo_last_error": (C.c_char_p, []),
    "s252_cairo_vm_run": (_i, [_vp, _sz, _u64, _u64, C.POINTER(_vp)]),
    "s252_cairo_vm_run_builtins": (_i, [_vp, _sz, _u64, _u64, C.c_uint, C.POINTER(_vp)]),
    "s252_cairo_run_segment": (_i, [_vp, _i, _vp]),
    "s252_cairo_run_destroy": (None, [_vp]),
    "s252_cairo_run_steps": (_sz, [_vp]),
    "s252_cairo_run_trace_len": (_sz, [_vp]),
    "s252_cairo_run_memory_len": (_sz, [_vp]),
    "s252_cairo_run_trace_bytes": (None, [_vp, _vp]),
    "s252_cairo_run_memory_bytes": (None, [_vp, _vp]),
    "s252_cairo_build_main_trace": (_i, [_vp, _sz, _vp, _sz, _sz, _vp, _vp, C.POINTER(_vp)]),
    "s252_cairo_build_execution_trace": (_i, [_vp, _sz, _vp, _sz, _sz, _vp, _vp, C.POINTER(_vp)]),
    "s252_cairo_trace_destroy": (None, [_vp]),
    "s252_cairo_trace_pin": (_i, [_vp]),
    "s252_cairo_trace_n_rows": (_sz, [_vp]),
    "s252_cairo_trace_n_cols": (_sz, [_vp]),
    "s252_cairo_trace_table": (_vp, [_vp]),
    "s252_cairo_trace_public_inputs": (None, [_vp, _vp]),
    "s252_cairo_trace_public_memory": (None, [_vp, _vp, _vp]),
    "s252_cairo_trace_serialize_public_inputs": (_sz, [_vp, _vp]),
    "s252_cairo_trace_from_table": (_i, [_vp, _sz, _sz, _vp, _vp, _vp, C.POINTER(_vp)]),
    "s252_cairo_round1": (_i, [_vp, _vp, _sz, _u64, _vp, C.POINTER(_vp), C.POINTER(_vp), _vp]),
    "s252_commit_read_trace": (_i, [_vp, _sz, _vp]),
    "s252_cairo_round2": (_i, [_vp, _vp, _vp, _vp, _vp, _sz, _u64, _vp, C.POINTER(_vp)]),
    "s252_cairo_constraint_evaluations": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _u64, _vp]),
    "s252_cairo_trace_columns": (_vp, [_vp]),
    "s252_lde_host_columns": (_i, [_vp, _vp, _sz, _sz, _sz, _u64, _i, C.POINTER(_vp)]),
    "s252_commit_device_trace": (_vp, [_vp]),
    "s252_lde_device_columns": (_i, [_vp, _vp, _sz, _sz, _sz, _u64, C.POINTER(_vp)]),
    "s252_cairo_aux_trace_device": (_i, [_vp, _vp, _vp, _vp, _i, C.POINTER(_vp)]),
    "s252_cairo_constraints_rows": (_i, [_vp, _vp, _vp, _vp, _sz, _sz, _sz, _vp, _vp, _sz, _vp, _vp, _vp, _sz, _u64, _vp]),
    "s252_cairo_composition_commit": (_i, [_vp, _vp, _sz, _sz, _u64, C.POINTER(_vp), _vp]),
    "s252_cairo_composition_lde": (_i, [_vp, _vp, _sz, _sz, _u64, C.POINTER(_vp)]),
    "s252_deep_rows": (_i, [_vp, _vp, _vp, _vp, _sz, _sz, _sz, _sz, _sz, _vp, _vp, _sz, _vp, _vp, _vp, _vp, _vp, _vp, _u64, _vp]),
    "s252_cairo_prove": (_i, [_vp, _vp, _sz, _sz, _u64, _u8, C.POINTER(_vp), C.POINTER(_sz)]),
    "s252_cairo_proof_free": (None, [_vp]),
    "s252_cairo_serialize_proof": (_i, [_sz, _vp, _vp, _vp, _vp, _sz, _vp, _sz, _vp, _vp, _sz, _sz, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp,
                                        _vp, _sz, _vp, _u64, C.POINTER(_vp), C.POINTER(_sz)]),
    "s252_cairo_last_prove_stages": (C.c_char_p, []),
}


def library_path():
    return _build.SO


def lib():
    """Loads the shared library (building it with nvcc if it is missing or stale)."""
    global _lib
    if _lib is None:
        path = _build.SO
        if not os.path.exists(path) or (_build.needs_build() and _build_available()):
            _build.build()
        L = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)   # AttributeError if the library lacks a declared symbol
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def _build_available():
    try:
        _build.nvcc_path()
        return True
    except RuntimeError:
        return False


class Stark252Error(RuntimeError):
    def __init__(self, code, message):
        super().__init__("%s (code %d)" % (message, code))
        self.code = code


class FFTError(Stark252Error):
    """Mirrors lambdaworks_math::fft::errors::FFTError (returned by the FFTPoly methods)."""


def ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return a.ctypes.data_as(C.c_void_p)


def fe_array(a, shape_tail=(4,)):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if a.shape[-1:] != shape_tail:
        raise ValueError("field elements are uint64[..., 4] (LW layout)")
    return a


class Context:
    """One GPU + one stream + twiddle cache (s252_ctx)."""

    def __init__(self, device=0):
        h = C.c_void_p()
        rc = lib().s252_ctx_create(device, C.byref(h))
        if rc != OK:
            raise Stark252Error(rc, "s252_ctx_create failed: no usable CUDA device %d (there is no CPU fallback)" % device)
        self.handle = h
        self.device = device
        self._children = weakref.WeakSet()    # live commit / FRI handles: they must not outlive the context

    def adopt(self, child):
        self._children.add(child)

    def close(self):
        if getattr(self, "handle", None):
            for child in list(self._children):
                child.free()
            lib().s252_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc, exc=Stark252Error):
        if rc != OK:
            raise exc(rc, lib().s252_last_error(self.handle).decode())

    def synchronize(self):
        self.check(lib().s252_ctx_synchronize(self.handle))

    def trim(self):
        """Return the arena's cached (unused) device blocks to the driver."""
        self.check(lib().s252_ctx_trim(self.handle))

    @property
    def stream(self):
        return lib().s252_ctx_stream(self.handle)

    @property
    def launch_count(self):
        return int(lib().s252_ctx_launch_count(self.handle))

    def profile(self, enable=True, reset=False):
        self.check(lib().s252_ctx_profile(self.handle, 2 if (enable and reset) else int(bool(enable))))

    def profile_read(self):
        import json
        buf = C.create_string_buffer(1 << 16)
        self.check(lib().s252_ctx_profile_read(self.handle, buf, len(buf)))
        return json.loads(buf.value.decode())

    # raw device buffers for S252_DEVICE calls (bench / multi-GPU plumbing)
    def device_alloc(self, nbytes):
        p = C.c_void_p()
        self.check(lib().s252_device_alloc(self.handle, nbytes, C.byref(p)))
        return p.value

    def device_free(self, p):
        self.check(lib().s252_device_free(self.handle, C.c_void_p(p)))

    def to_device(self, dptr, arr):
        arr = np.ascontiguousarray(arr)
        self.check(lib().s252_copy_to_device(self.handle, C.c_void_p(dptr), ptr(arr), arr.nbytes))

    def to_device_async(self, dptr, host_ptr, nbytes):
        """Prefetch on the copy stream (host_ptr: address of pinned host memory)."""
        self.check(lib().s252_copy_to_device_async(self.handle, C.c_void_p(dptr), C.c_void_p(host_ptr), nbytes))

    def copy_stream_wait(self):
        self.check(lib().s252_copy_stream_wait(self.handle))

    def to_host(self, arr, dptr):
        self.check(lib().s252_copy_to_host(self.handle, ptr(arr), C.c_void_p(dptr), arr.nbytes))


_default_ctx = {}


def default_context(device=None):
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0")) if os.environ.get("S252_USE_LOCAL_RANK") else 0
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]
