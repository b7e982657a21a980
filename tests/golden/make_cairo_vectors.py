"""Extracts the expected execution traces held by the reference's own tests
(src/cairo/execution_trace.rs: test_build_main_trace_simple_program, ..._call_func_program) into
tests/golden/cairo/expected_traces.json.  Run in the build container (reads /root/reference); the
JSON travels with the repo.

The compiled programs themselves are git-ignored in the reference, so each program's bytecode is
recovered from the expected trace: mem[pc] = the `inst` column, and for immediate operands
mem[op1_addr] = the `op1` column.
"""
import json
import os
import re

SRC = "/root/reference/src/cairo/execution_trace.rs"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cairo", "expected_traces.json")
P = 2**251 + 17 * 2**192 + 1


def parse_cols(body):
    cols = []
    for chunk in re.split(r"//\s*col \d+[^\n]*\n", body)[1:]:
        vals = []
        for m in re.finditer(r"FE::zero\(\)|FE::one\(\)|FE::from\(\s*(0x[0-9a-fA-F]+|\d+)\s*\)|"
                             r"FE::from_hex_unchecked\(\s*\"([0-9a-fA-F]+)\",?\s*\)", chunk):
            t = m.group(0)
            if t.startswith("FE::zero"):
                vals.append(0)
            elif t.startswith("FE::one"):
                vals.append(1)
            elif m.group(1):
                vals.append(int(m.group(1), 0))
            else:
                vals.append(int(m.group(2), 16))
        cols.append(vals)
    return cols


def main():
    src = open(SRC).read()
    out = {}
    for name in ("test_build_main_trace_simple_program", "test_build_main_trace_call_func_program"):
        start = src.index("fn " + name)
        a = src.index("let expected_trace = TraceTable::new_from_cols(", start)
        b = src.index("assert_eq!(execution_trace.cols(), expected_trace.cols());", a)
        cols = parse_cols(src[a:b])
        assert len(cols) == 34 and len({len(c) for c in cols}) == 1, (name, len(cols))
        pcs, insts, op1_addrs, op1s, imm = cols[19], cols[23], cols[22], cols[26], cols[2]
        program_size = cols[17][0] - 3          # ap_0 = 1 + program_size + 2
        mem = {}
        for i in range(len(pcs)):
            mem[pcs[i]] = insts[i]
            if imm[i]:
                mem[op1_addrs[i]] = op1s[i]
        program = [mem.get(a) for a in range(1, program_size + 1)]
        out[name] = {"columns": [[hex(v) for v in c] for c in cols], "program": [None if w is None else hex(w) for w in program],
                     "source": "src/cairo/execution_trace.rs"}
    # cairo-run dumps of the "mul" program held by the reference's tests (register_states.rs:95-115,
    # cairo_mem.rs:74-95): a known answer for the machine's relocated trace and memory files
    regs_src = open("/root/reference/src/cairo/register_states.rs").read()
    mem_src = open("/root/reference/src/cairo/cairo_mem.rs").read()
    a = regs_src.index("fn mul_program_gives_expected_trace")
    trace_hex = re.search(r'hex::decode\("([0-9a-f]+)"\)', regs_src[a:]).group(1)
    a = mem_src.index("fn mem_indexes_are_contiguos_in_bytes_of_mul_program")
    mem_hex = re.search(r'hex::decode\("([0-9a-f]+)"\)', mem_src[a:]).group(1)
    out["mul_program_cairo_run_dump"] = {"trace_hex": trace_hex, "memory_hex": mem_hex,
                                         "source": "src/cairo/register_states.rs, src/cairo/cairo_mem.rs"}
    json.dump(out, open(OUT, "w"), indent=0)
    print(OUT, sorted(out))


if __name__ == "__main__":
    main()
