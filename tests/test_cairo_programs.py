"""CPU: hand-assembled Cairo-0 programs that exercise every instruction flavour the machine and the trace
builder implement (immediates, ap/fp operands, add/mul, jnz taken and not taken, call/ret, jmp rel/abs,
ap += imm with the memory holes it leaves).  There is no expected trace for them in the reference; the
check is the AIR itself: the oracle's prover only succeeds (composition polynomial within its degree
bound) when all 49 transition constraints and the 8 boundary constraints hold on the trace the
library built, and the oracle's verifier (pinned on the reference's golden proof) must accept."""
import numpy as np
import pytest

from lambdaworks_cairo_prover_b200 import ProofOptions, cairo, felt
from oracle.cairo_prover import cairo_prove
from oracle.cairo_verifier import cairo_verify
from oracle.proof_format import StarkProof

P = felt.MODULUS
OP1 = {"op0": 0, "imm": 1, "fp": 2, "ap": 4}
RES = {"op1": 0, "add": 1, "mul": 2}
PC = {"regular": 0, "abs": 1, "rel": 2, "jnz": 4}
AP = {"regular": 0, "add": 1, "add1": 2}
OPC = {"nop": 0, "call": 1, "ret": 2, "assert_eq": 4}
RET = 0x208b7fff7fff7ffe


def ins(off_dst, off_op0, off_op1, dst_fp=0, op0_fp=1, op1="op0", res="op1", pc="regular", ap="regular", opcode="nop"):
    """Cairo whitepaper section 4.5: three biased 16-bit offsets and 15 flag bits."""
    flags = dst_fp | op0_fp << 1 | OP1[op1] << 2 | RES[res] << 5 | PC[pc] << 7 | AP[ap] << 10 | OPC[opcode] << 12
    return (off_dst + 0x8000) | (off_op0 + 0x8000) << 16 | (off_op1 + 0x8000) << 32 | flags << 48


def push_imm(v):        # [ap] = v; ap++
    return [ins(0, -1, 1, op1="imm", ap="add1", opcode="assert_eq"), v % P]


def test_assembler_reproduces_known_encodings():
    assert push_imm(1)[0] == 0x480680017fff8000
    assert ins(-2, -1, -1, dst_fp=1, op0_fp=1, op1="fp", pc="abs", opcode="ret") == RET
    assert ins(0, 1, 1, op0_fp=0, op1="imm", pc="rel", ap="regular", opcode="call") == 0x1104800180018000
    assert ins(-3, -1, 1, dst_fp=1, op1="imm", pc="jnz") == 0x20780017fff7ffd
    assert ins(0, -5, -4, op0_fp=1, op1="fp", res="add", ap="add1", opcode="assert_eq") == 0x482a7ffc7ffb8000


PROGRAMS = {
    # immediates, ap- and fp-relative operands, add and mul
    "arith": push_imm(7) + push_imm(9) + [
        ins(0, -1, -2, op0_fp=0, op1="ap", res="mul", ap="add1", opcode="assert_eq"),          # [ap] = [ap-1] * [ap-2]; ap++
        ins(0, -1, 0, op0_fp=0, op1="fp", res="add", ap="add1", opcode="assert_eq"),           # [ap] = [ap-1] + [fp]; ap++
        ins(0, -1, 1, op0_fp=0, op1="imm", res="add", ap="add1", opcode="assert_eq"), 12345,   # [ap] = [ap-1] + 12345; ap++
        ins(0, -1, 1, op0_fp=0, op1="imm", res="mul", ap="add1", opcode="assert_eq"), P - 3,   # [ap] = [ap-1] * (-3); ap++
        RET],
    # a countdown: jnz taken four times, then not taken
    "loop": push_imm(4) + [
        ins(0, -1, 1, op0_fp=0, op1="imm", res="add", ap="add1", opcode="assert_eq"), P - 1,   # loop: [ap] = [ap-1] - 1; ap++
        ins(-1, -1, 1, dst_fp=0, op1="imm", pc="jnz"), P - 2,                                  # jmp rel -2 if [ap-1] != 0
        RET],
    # call / ret with an argument read through fp
    "call": push_imm(3) + [
        ins(0, 1, 1, op0_fp=0, op1="imm", pc="rel", opcode="call"), 5,                        # call rel 5 -> f
        ins(0, -1, 1, op0_fp=0, op1="imm", res="add", ap="add1", opcode="assert_eq"), 1,      # [ap] = [ap-1] + 1; ap++
        RET,
        ins(0, -3, -3, op0_fp=1, op1="fp", res="mul", ap="add1", opcode="assert_eq"),         # f: [ap] = [fp-3] * [fp-3]; ap++
        RET],
    # jmp rel over an instruction, ap += 2 (leaves two unwritten cells: memory holes), jmp abs
    "jumps": [
        ins(-1, -1, 1, dst_fp=1, op1="imm", pc="rel"), 4,                                      # jmp rel 4
        ] + push_imm(999) + push_imm(1) + [
        ins(-1, -1, 1, dst_fp=1, op1="imm", ap="add"), 2,                                      # ap += 2
        ] + push_imm(5) + [
        ins(-1, -1, 1, dst_fp=1, op1="imm", pc="abs"), 1 + 14,                                 # jmp abs -> the ret below
        ] + push_imm(777) + [
        RET],
}


@pytest.mark.parametrize("name", sorted(PROGRAMS))
def test_program_trace_satisfies_the_air(name):
    words = PROGRAMS[name]
    regs, mem, size = cairo.run_program(words)
    t = cairo.build_main_trace(regs, mem, size)
    steps = len(regs) // 24
    assert steps == {"arith": 7, "loop": 10, "call": 6, "jumps": 6}[name]
    table = np.array(t.table).reshape(t.n_rows(), t.n_cols, 4)
    opts = ProofOptions.default_test_options()
    proof = cairo_prove(table, t.pub_inputs, opts, threads=1)              # raises if a constraint is violated
    assert cairo_verify(StarkProof.parse(proof.serialize()), t.pub_inputs, opts)


def test_expected_values_of_the_arith_program():
    regs, mem, size = cairo.run_program(PROGRAMS["arith"])
    cells = {int.from_bytes(mem[40 * i:40 * i + 8], "little"): int.from_bytes(mem[40 * i + 8:40 * i + 40], "little") for i in range(len(mem) // 40)}
    base = size + 3                                                          # first free cell: program, return fp, return pc
    assert [cells[base + k] for k in range(6)] == [7, 9, 63, 70, 12415, (-3 * 12415) % P]


# ---------------------------------------------------------------------------------------------------------------
# Programs that declare builtins (hint-free equivalents of the reference's cairo_programs/cairo0/rc_program.cairo and
# output_program.cairo / signed_div_rem.cairo, tests/integration_tests.rs:153-172: those use library functions whose hints
# need cairo-lang; here the range-checked / output values are written directly).  main's frame: the builtin pointers sit
# below [return_fp, return_pc], output first.  The 43-column layout (range-check builtin columns, air.rs:594-625) comes out
# of a real execution.
def assert_at_ptr(ptr_fp_off, k):      # [[fp + ptr_fp_off] + k] = [ap - 1]
    return ins(-1, ptr_fp_off, k, dst_fp=0, op0_fp=1, op1="op0", opcode="assert_eq")


def ret_ptr_plus(ptr_fp_off, k):       # [ap] = [fp + ptr_fp_off] + k; ap++
    return [ins(0, ptr_fp_off, 1, dst_fp=0, op0_fp=1, op1="imm", res="add", ap="add1", opcode="assert_eq"), k]


BUILTIN_PROGRAMS = {
    # %builtins range_check: three values through the builtin (assert_nn(5), assert_nn(2) in rc_program.cairo range-check 5 and 2)
    "rc": (("range_check",),
           push_imm(5) + [assert_at_ptr(-3, 0)] + push_imm(2) + [assert_at_ptr(-3, 1)] + push_imm(2**100 + 12345) + [assert_at_ptr(-3, 2)]
           + ret_ptr_plus(-3, 3) + [RET]),
    # %builtins output range_check: serialize_word twice, one range check, both pointers returned
    "output_rc": (("output", "range_check"),
                  push_imm(1234) + [assert_at_ptr(-4, 0)] + push_imm(P - 4) + [assert_at_ptr(-4, 1)] + push_imm(70000) + [assert_at_ptr(-3, 0)]
                  + ret_ptr_plus(-4, 2) + ret_ptr_plus(-3, 1) + [RET]),
    # a loop that range-checks a countdown: the pointer is threaded through ap like compiled code does
    "rc_loop": (("range_check",),
                [ins(0, -3, -3, dst_fp=0, op0_fp=1, op1="fp", ap="add1", opcode="assert_eq")]      # [ap] = [fp-3] (rc ptr); ap++
                + push_imm(3)                                                                       # counter
                + [ins(-1, -2, 0, dst_fp=0, op0_fp=0, op1="op0", opcode="assert_eq"),              # loop: [[ap-2]] = [ap-1]
                   ins(0, -2, 1, dst_fp=0, op0_fp=0, op1="imm", res="add", ap="add1", opcode="assert_eq"), 1,       # [ap] = [ap-2] + 1 (ptr)
                   ins(0, -2, 1, dst_fp=0, op0_fp=0, op1="imm", res="add", ap="add1", opcode="assert_eq"), P - 1,   # [ap] = [ap-2] - 1 (counter)
                   ins(-1, -1, 1, dst_fp=0, op1="imm", pc="jnz"), P - 5,                           # jmp loop if counter != 0
                   ins(0, -2, -2, dst_fp=0, op0_fp=0, op1="ap", ap="add1", opcode="assert_eq"),    # [ap] = [ap-2] (final ptr); ap++
                   RET]),
}


@pytest.mark.parametrize("name", sorted(BUILTIN_PROGRAMS))
def test_builtin_program_trace_satisfies_the_air(name):
    builtins, words = BUILTIN_PROGRAMS[name]
    regs, mem, size, rc_range, out_range = cairo.run_program_with_builtins(words, builtins)
    assert (rc_range is not None) == ("range_check" in builtins) and (out_range is not None) == ("output" in builtins)
    cells = {int.from_bytes(mem[40 * i:40 * i + 8], "little"): int.from_bytes(mem[40 * i + 8:40 * i + 40], "little") for i in range(len(mem) // 40)}
    if name == "rc":
        assert [cells[a] for a in range(*rc_range)] == [5, 2, 2**100 + 12345]
    if name == "output_rc":
        assert [cells[a] for a in range(*out_range)] == [1234, P - 4] and [cells[a] for a in range(*rc_range)] == [70000]
        assert out_range[1] == rc_range[0]                        # segments are laid out in declaration order
    if name == "rc_loop":
        assert [cells[a] for a in range(*rc_range)] == [3, 2, 1]
    t = cairo.build_main_trace(regs, mem, size, rc_range, out_range)
    assert t.n_cols == 43                                         # 34 + the nine range-check builtin columns
    table = np.array(t.table).reshape(t.n_rows(), t.n_cols, 4)
    opts = ProofOptions.default_test_options()
    proof = cairo_prove(table, t.pub_inputs, opts, threads=1)     # raises if one of the 50 transition constraints is violated
    assert cairo_verify(StarkProof.parse(proof.serialize()), t.pub_inputs, opts)


def test_range_check_builtin_rejects_wide_values():
    words = push_imm(2**128) + [assert_at_ptr(-3, 0)] + ret_ptr_plus(-3, 1) + [RET]
    with pytest.raises(Exception):
        cairo.run_program_with_builtins(words, ("range_check",))
