"""Cairo side of the prover: the reference's `cairo::runner::run` and `cairo::execution_trace` interface
over include/stark252_cairo.h (host code in the shared library, C++).

    run_program            src/cairo/runner/run.rs:62-241   (a minimal Cairo-0 machine replaces cairo-vm)
    generate_prover_args   src/cairo/runner/run.rs:243-266
    build_main_trace       src/cairo/execution_trace.rs:57-87
    PublicInputs           src/cairo/air.rs:155-276
"""
import ctypes as C
import json

import numpy as np

from . import _native as N
from .merkle import DeviceCommit
from .prover import TraceTable


class CairoError(RuntimeError):
    pass


def _check(rc):
    if rc != N.OK:
        raise CairoError(N.lib().s252_cairo_last_error().decode())


class _PublicInputsC(C.Structure):
    _fields_ = [("pc_init", C.c_uint64), ("ap_init", C.c_uint64), ("fp_init", C.c_uint64), ("pc_final", C.c_uint64),
                ("ap_final", C.c_uint64), ("num_steps", C.c_uint64), ("n_public_memory", C.c_uint64),
                ("rc_segment", C.c_uint64 * 2), ("output_segment", C.c_uint64 * 2),
                ("range_check_min", C.c_uint16), ("range_check_max", C.c_uint16),
                ("has_range_check_bounds", C.c_uint8), ("has_rc_segment", C.c_uint8), ("has_output_segment", C.c_uint8),
                ("reserved", C.c_uint8)]


class PublicInputs:
    """src/cairo/air.rs:155-176.  `public_memory` is {address: LW element}."""

    def __init__(self, c, addrs, values, serialized):
        self.pc_init, self.ap_init, self.fp_init = c.pc_init, c.ap_init, c.fp_init
        self.pc_final, self.ap_final, self.num_steps = c.pc_final, c.ap_final, c.num_steps
        self.range_check_min = c.range_check_min if c.has_range_check_bounds else None
        self.range_check_max = c.range_check_max if c.has_range_check_bounds else None
        self.memory_segments = {}
        if c.has_rc_segment:
            self.memory_segments["RangeCheck"] = range(c.rc_segment[0], c.rc_segment[1])
        if c.has_output_segment:
            self.memory_segments["Output"] = range(c.output_segment[0], c.output_segment[1])
        self.public_memory = {int(a): v for a, v in zip(addrs, values)}
        self._serialized = serialized

    def serialize(self):
        """PublicInputs::serialize (air.rs:217-276), public memory in address order."""
        return self._serialized


class MainTrace(TraceTable):
    """The table build_main_trace returns, plus the handle the GPU prover consumes."""

    def __init__(self, handle):
        L = N.lib()
        self.handle = handle
        n_rows, n_cols = L.s252_cairo_trace_n_rows(handle), L.s252_cairo_trace_n_cols(handle)
        base = L.s252_cairo_trace_table(handle)
        buf = (C.c_uint64 * (n_rows * n_cols * 4)).from_address(base)
        self.table = np.frombuffer(buf, dtype=np.uint64).reshape(-1, 4)     # zero-copy view of the library's table
        self.n_cols = n_cols
        c = _PublicInputsC()
        L.s252_cairo_trace_public_inputs(handle, C.byref(c))
        addrs = np.zeros(c.n_public_memory, dtype=np.uint64)
        values = np.zeros((c.n_public_memory, 4), dtype=np.uint64)
        L.s252_cairo_trace_public_memory(handle, N.ptr(addrs), N.ptr(values))
        ser = np.zeros(L.s252_cairo_trace_serialize_public_inputs(handle, None), dtype=np.uint8)
        L.s252_cairo_trace_serialize_public_inputs(handle, N.ptr(ser))
        self.pub_inputs = PublicInputs(c, addrs, values, ser.tobytes())

    def free(self):
        if getattr(self, "handle", None):
            self.table = None
            N.lib().s252_cairo_trace_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def program_words(program_content):
    """The `data` array of a compiled Cairo-0 program JSON -> list of ints."""
    prog = json.loads(program_content) if isinstance(program_content, (bytes, str)) else program_content
    return [int(x, 16) for x in prog["data"]]


def run_program(words, entry_offset=0, max_steps=0):
    """run_program(None, layout, program, V0) (run.rs:62-241) for hint-free, builtin-free programs.
    words: program data as ints.  Returns (register_states_bytes, memory_bytes, program_size) in the
    reference's binary formats (register_states.rs:47-78, cairo_mem.rs:35-61)."""
    L = N.lib()
    be = b"".join(int(w).to_bytes(32, "big") for w in words)
    buf = np.frombuffer(be, dtype=np.uint8)
    h = C.c_void_p()
    _check(L.s252_cairo_vm_run(N.ptr(buf), len(words), entry_offset, max_steps, C.byref(h)))
    try:
        trace = np.zeros(L.s252_cairo_run_trace_len(h), dtype=np.uint8)
        memory = np.zeros(L.s252_cairo_run_memory_len(h), dtype=np.uint8)
        L.s252_cairo_run_trace_bytes(h, N.ptr(trace))
        L.s252_cairo_run_memory_bytes(h, N.ptr(memory))
    finally:
        L.s252_cairo_run_destroy(h)
    return trace.tobytes(), memory.tobytes(), len(words)


BUILTINS = {"output": 1, "range_check": 2}


def run_program_with_builtins(words, builtins, entry_offset=0, max_steps=0):
    """run_program for a program that declares builtins ("output", "range_check": cairo-vm's BuiltinRunner for the two the AIR
    knows).  Returns (register_states_bytes, memory_bytes, program_size, rc_range, output_range): the ranges are what
    generate_prover_args hands to build_main_trace (run.rs:243-266), None for an absent builtin."""
    L = N.lib()
    be = b"".join(int(w).to_bytes(32, "big") for w in words)
    buf = np.frombuffer(be, dtype=np.uint8)
    mask = 0
    for b in builtins:
        mask |= BUILTINS[b]
    h = C.c_void_p()
    _check(L.s252_cairo_vm_run_builtins(N.ptr(buf), len(words), entry_offset, max_steps, mask, C.byref(h)))
    try:
        trace = np.zeros(L.s252_cairo_run_trace_len(h), dtype=np.uint8)
        memory = np.zeros(L.s252_cairo_run_memory_len(h), dtype=np.uint8)
        L.s252_cairo_run_trace_bytes(h, N.ptr(trace))
        L.s252_cairo_run_memory_bytes(h, N.ptr(memory))
        ranges = []
        for which in (0, 1):
            r = np.zeros(2, dtype=np.uint64)
            present = L.s252_cairo_run_segment(h, which, N.ptr(r))
            ranges.append((int(r[0]), int(r[1])) if present else None)
    finally:
        L.s252_cairo_run_destroy(h)
    return trace.tobytes(), memory.tobytes(), len(words), ranges[0], ranges[1]


def _range_arg(r):
    return None if r is None else np.array([r[0], r[1]] if not isinstance(r, range) else [r.start, r.stop], dtype=np.uint64)


def build_main_trace(register_states, memory, program_size, rc_range=None, output_range=None, execution_only=False):
    """PublicInputs::from_regs_and_mem + build_main_trace (execution_trace.rs:57-87).  Returns a
    MainTrace whose .pub_inputs carry range_check_min/max.  execution_only: build_cairo_execution_trace
    (execution_trace.rs:261-356) without holes, dummy accesses and padding."""
    L = N.lib()
    t = np.frombuffer(register_states, dtype=np.uint8)
    m = np.frombuffer(memory, dtype=np.uint8)
    rc, out = _range_arg(rc_range), _range_arg(output_range)
    h = C.c_void_p()
    fn = L.s252_cairo_build_execution_trace if execution_only else L.s252_cairo_build_main_trace
    _check(fn(N.ptr(t), t.size, N.ptr(m), m.size, program_size, N.ptr(rc), N.ptr(out), C.byref(h)))
    return MainTrace(h)


def generate_prover_args(program_content, output_range=None, entry_offset=0):
    """run.rs:243-266: (main_trace, pub_inputs) for a compiled Cairo-0 program."""
    words = program_words(program_content)
    regs, mem, size = run_program(words, entry_offset)
    trace = build_main_trace(regs, mem, size, None, output_range)
    return trace, trace.pub_inputs


def fibonacci_program(n, with_assert=None):
    """Compiled `cairo_programs/cairo0/fibonacci_N.cairo` (main calls fib(1, 1, n), optionally
    asserts the result): the 22/24-word bytecode is the public memory of the reference's proofs
    (benches/proofs/*.proof); only the immediate n (and the asserted value) change."""
    P = 2**251 + 17 * 2**192 + 1
    main = [0x480680017fff8000, 1, 0x480680017fff8000, 1, 0x480680017fff8000, n, 0x1104800180018000]
    fib = [0x20780017fff7ffd, 5, 0x480a7ffc7fff8000, 0x480a7ffc7fff8000, 0x208b7fff7fff7ffe,
           0x482a7ffc7ffb8000, 0x480a7ffc7fff8000, 0x48127ffe7fff8000, 0x482680017ffd8000, P - 1,
           0x1104800180018000, P - 10, 0x208b7fff7fff7ffe]
    if with_assert is None:
        return main + [3, 0x208b7fff7fff7ffe] + fib
    return main + [5, 0x400680017fff7fff, with_assert, 0x208b7fff7fff7ffe] + fib


# ------------------------------------------------------------------------------------------ GPU prover
def trace_from_table(table, n_cols, pub_inputs):
    """A MainTrace around a caller-built table (row-major LW) and PublicInputs-like object
    (s252_cairo_trace_from_table): what a Rust caller holding TraceTable + PublicInputs would pass."""
    L = N.lib()
    table = N.fe_array(np.asarray(table, dtype=np.uint64).reshape(-1, 4))
    c = _PublicInputsC()
    for f in ("pc_init", "ap_init", "fp_init", "pc_final", "ap_final", "num_steps"):
        setattr(c, f, getattr(pub_inputs, f))
    if pub_inputs.range_check_min is not None:
        c.has_range_check_bounds, c.range_check_min, c.range_check_max = 1, pub_inputs.range_check_min, pub_inputs.range_check_max
    for name, flag, field in (("RangeCheck", "has_rc_segment", "rc_segment"), ("Output", "has_output_segment", "output_segment")):
        if name in pub_inputs.memory_segments:
            r = pub_inputs.memory_segments[name]
            setattr(c, flag, 1)
            getattr(c, field)[0], getattr(c, field)[1] = r.start, r.stop
    addrs = np.array(sorted(pub_inputs.public_memory), dtype=np.uint64)
    vals = np.stack([pub_inputs.public_memory[int(a)] for a in addrs]) if len(addrs) else np.zeros((0, 4), dtype=np.uint64)
    c.n_public_memory = len(addrs)
    h = C.c_void_p()
    _check(L.s252_cairo_trace_from_table(N.ptr(table), table.shape[0] // n_cols, n_cols, C.byref(c), N.ptr(addrs),
                                         N.ptr(N.fe_array(vals)), C.byref(h)))
    return MainTrace(h)


class Round1Commit(DeviceCommit):
    """A round-1 handle that also keeps the trace evaluations it was built from."""

    def trace_column(self, col):
        out = np.empty((self.n_coeffs, 4), dtype=np.uint64)
        self.ctx.check(N.lib().s252_commit_read_trace(self.handle, col, N.ptr(out)))
        return out


def round_1_randomized_air_with_preprocessing(trace, options, transcript, ctx=None):
    """src/starks/prover.rs:186-224 for CairoAIR on the GPU (auxiliary trace built on the device).
    Returns (main_commit, aux_commit, rap_challenges[3, 4])."""
    ctx = ctx or N.default_context()
    hm, ha = C.c_void_p(), C.c_void_p()
    rap = np.zeros((3, 4), dtype=np.uint64)
    ctx.check(N.lib().s252_cairo_round1(ctx.handle, trace.handle, options.blowup_factor, options.coset_offset, transcript.handle,
                                        C.byref(hm), C.byref(ha), N.ptr(rap)))
    mk = lambda h: Round1Commit(ctx, h, _root_of(ctx, h))
    return mk(hm), mk(ha), rap


def _root_of(ctx, h):
    root = np.empty(32, dtype=np.uint8)
    ctx.check(N.lib().s252_commit_root(h, N.ptr(root)))
    return root.tobytes()


def round_2_compute_composition_polynomial(trace, main_commit, aux_commit, rap, options, transcript):
    """prover.rs:598-640 + 226-283 for CairoAIR on the GPU: samples the coefficients, evaluates the
    constraints over the LDE coset, commits H1/H2 and appends the root.  Returns the composition commit."""
    ctx = main_commit.ctx
    h = C.c_void_p()
    ctx.check(N.lib().s252_cairo_round2(ctx.handle, trace.handle, main_commit.handle, aux_commit.handle, N.ptr(N.fe_array(rap)),
                                        options.blowup_factor, options.coset_offset, transcript.handle, C.byref(h)))
    return DeviceCommit(ctx, h, _root_of(ctx, h))


def constraint_evaluations(trace, main_commit, aux_commit, rap, boundary_coeffs, transition_coeffs, options):
    """ConstraintEvaluator::evaluate (constraints/evaluator.rs:40-262) for CairoAIR on the GPU with
    explicit coefficients: boundary_coeffs [8, 2, 4], transition_coeffs [49|50, 2, 4] as (alpha, beta).
    Returns the evaluations of the composition polynomial on the LDE coset, uint64[M, 4]."""
    ctx = main_commit.ctx
    out = np.empty((main_commit.n_rows, 4), dtype=np.uint64)
    ctx.check(N.lib().s252_cairo_constraint_evaluations(ctx.handle, trace.handle, main_commit.handle, aux_commit.handle,
                                                        N.ptr(N.fe_array(rap)), N.ptr(N.fe_array(boundary_coeffs)),
                                                        N.ptr(N.fe_array(transition_coeffs)), options.blowup_factor,
                                                        options.coset_offset, N.ptr(out)))
    return out


def generate_cairo_proof(trace, proof_options, ctx=None):
    """src/cairo/air.rs:1183-1190: prove::<Stark252PrimeField, CairoAIR>(trace, pub_inputs, options) on
    the GPU.  Returns StarkProof::serialize() bytes."""
    ctx = ctx or N.default_context()
    out, n = C.c_void_p(), C.c_size_t()
    ctx.check(N.lib().s252_cairo_prove(ctx.handle, trace.handle, proof_options.blowup_factor, proof_options.fri_number_of_queries,
                                       proof_options.coset_offset, proof_options.grinding_factor, C.byref(out), C.byref(n)))
    try:
        return C.string_at(out.value, n.value)
    finally:
        N.lib().s252_cairo_proof_free(out)
