// cairo_host.hpp -- host side of the Cairo workload: a minimal Cairo-0 machine and the main-trace
// builder (the callers of the LDE + commit path).
//
//   * Vm            stands in for cairo-vm 0.6.0 as the reference drives it (src/cairo/runner/run.rs:
//                   62-100: layout "small", proof_mode = false, relocate_mem = true, no hints, no
//                   builtins): it executes plain Cairo instructions (Cairo whitepaper section 4.5)
//                   and emits the *relocated* register trace and memory in the reference's own
//                   binary formats (register_states.rs:47-78: 24 B LE rows ap,fp,pc;
//                   cairo_mem.rs:35-61: 40 B LE rows address,value).
//   * build_main_trace   src/cairo/execution_trace.rs:57-87 and everything it calls.
//
// Values are Montgomery `fe`s (internal limb order); host arithmetic from host_field.hpp.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <exception>
#include <thread>
#include <unordered_map>
#include <vector>

#include "host_field.hpp"

namespace s252 {
namespace cairo {

namespace H = s252::host;

// fn(lo, hi, worker) over [0, n) on the host's cores (the front-end is outside the prover's timed region, but a
// 2^19-row trace is 18 M field elements: the row loop, the format conversion and the transposition are threaded)
template <typename F>
inline void parallel_for(size_t n, F fn) {
    unsigned nt = std::thread::hardware_concurrency();
    if (nt == 0) nt = 1;
    if (nt > 32) nt = 32;
    if (n < 4096 || nt == 1) { fn((size_t)0, n, 0u); return; }
    // an exception inside a worker (std::bad_alloc from a push_back, say) must not end in std::terminate: it is carried
    // to the calling thread and rethrown there, where the C entry points catch it
    std::vector<std::thread> th;
    std::vector<std::exception_ptr> errs(nt);
    for (unsigned t = 0; t < nt; ++t)
        th.emplace_back([=, &errs] {
            try { fn(n * t / nt, n * (t + 1) / nt, t); } catch (...) { errs[t] = std::current_exception(); }
        });
    for (auto& x : th) x.join();
    for (auto& e : errs) if (e) std::rethrow_exception(e);
}

// Column indices of the main trace (src/cairo/air.rs:74-125)
enum : unsigned {
    F_DST_FP = 0, F_OP_0_FP = 1, F_OP_1_VAL = 2, F_OP_1_FP = 3, F_OP_1_AP = 4, F_RES_ADD = 5, F_RES_MUL = 6,
    F_PC_ABS = 7, F_PC_REL = 8, F_PC_JNZ = 9, F_AP_ADD = 10, F_AP_ONE = 11, F_OPC_CALL = 12, F_OPC_RET = 13,
    F_OPC_AEQ = 14,
    FRAME_RES = 16, FRAME_AP = 17, FRAME_FP = 18, FRAME_PC = 19, FRAME_DST_ADDR = 20, FRAME_OP0_ADDR = 21,
    FRAME_OP1_ADDR = 22, FRAME_INST = 23, FRAME_DST = 24, FRAME_OP0 = 25, FRAME_OP1 = 26, OFF_DST = 27, OFF_OP0 = 28,
    OFF_OP1 = 29, FRAME_T0 = 30, FRAME_T1 = 31, FRAME_MUL = 32, FRAME_SELECTOR = 33,
    MAIN_COLS = 34, RC_BUILTIN_COLS = 9,
};

inline uint64_t low64(const fe& mont) {   // aux_get_last_nim_of_field_element, decode/instruction_flags.rs:18-31
    return H::to_u256(H::from_mont(mont)).w[0];
}
inline fe fe_from_le32(const uint8_t* b) {   // FE::from_bytes_le
    H::U256 c;
    for (int i = 0; i < 4; ++i) {
        uint64_t v = 0;
        for (int k = 7; k >= 0; --k) v = (v << 8) | b[8 * i + k];
        c.w[i] = v;
    }
    return H::to_mont(H::from_u256(c));
}
inline void fe_to_le32(const fe& mont, uint8_t* b) {
    const H::U256 c = H::to_u256(H::from_mont(mont));
    for (int i = 0; i < 4; ++i)
        for (int k = 0; k < 8; ++k) b[8 * i + k] = (uint8_t)(c.w[i] >> (8 * k));
}

// Decoded instruction (decode/instruction_flags.rs, decode/instruction_offsets.rs)
struct Instr {
    uint16_t flags;                 // 15 flag bits, bit i = trace column i
    uint16_t off[3];                // biased offsets off_dst, off_op0, off_op1 (trace representation)
    int32_t soff[3];                // signed offsets
    bool bit(unsigned i) const { return (flags >> i) & 1; }
    unsigned op1_src() const { return (flags >> 2) & 7; }     // 0 op0, 1 imm, 2 fp, 4 ap
    unsigned res_logic() const { return (flags >> 5) & 3; }   // 0 op1, 1 add, 2 mul
    unsigned pc_update() const { return (flags >> 7) & 7; }   // 0 regular, 1 abs, 2 rel, 4 jnz
    unsigned ap_update() const { return (flags >> 10) & 3; }  // 0 regular, 1 add, 2 add1
    unsigned opcode() const { return (flags >> 12) & 7; }     // 0 nop, 1 call, 2 ret, 4 assert_eq
    unsigned size() const { return op1_src() == 1 ? 2 : 1; }
};
inline bool decode(const fe& word, Instr* out, std::string* err) {
    const uint64_t w = low64(word);
    Instr in;
    for (int k = 0; k < 3; ++k) {
        in.off[k] = (uint16_t)(w >> (16 * k));
        in.soff[k] = (int32_t)in.off[k] - 0x8000;
    }
    in.flags = (uint16_t)((w >> 48) & 0x7fff);
    const unsigned s = in.op1_src(), r = in.res_logic(), p = in.pc_update(), a = in.ap_update(), o = in.opcode();
    if (!(s == 0 || s == 1 || s == 2 || s == 4)) { if (err) *err = "InvalidOp1Src"; return false; }
    if (r > 2) { if (err) *err = "InvalidResLogic"; return false; }
    if (!(p == 0 || p == 1 || p == 2 || p == 4)) { if (err) *err = "InvalidPcUpdate"; return false; }
    if (a > 2) { if (err) *err = "InvalidApUpdate"; return false; }
    if (!(o == 0 || o == 1 || o == 2 || o == 4)) { if (err) *err = "InvalidOpcode"; return false; }
    *out = in;
    return true;
}

struct RegisterState { uint64_t pc, fp, ap; };
struct Memory {
    std::unordered_map<uint64_t, fe> data;
    const fe* get(uint64_t a) const { auto it = data.find(a); return it == data.end() ? nullptr : &it->second; }
};

// ---------------------------------------------------------------------------------------------
// The machine.  Program at addresses 1..n, execution segment right after it; main is entered with
// the stack [return_fp, end] (two empty segments, relocated after the run) and the run stops when
// the final `ret` jumps to `end`.
// Builtins (cairo-vm's BuiltinRunner, as far as hint-free programs need them): `output` and `range_check`, each a memory
// segment of its own whose base pointer main receives on the stack in the order of the %builtins directive
// (output before range_check); after the run the segments are laid out behind the execution segment
// (run.rs:62-100 relocates; generate_prover_args, run.rs:243-266, hands the range-check range to the AIR).
enum : unsigned { BUILTIN_OUTPUT = 1, BUILTIN_RANGE_CHECK = 2 };
struct VmResult {
    std::vector<RegisterState> trace;
    std::vector<std::pair<uint64_t, fe>> memory;   // sorted by address
    size_t program_size = 0;
    bool has_output = false, has_rc = false;
    uint64_t output_range[2] = {0, 0}, rc_range[2] = {0, 0};   // [begin, end) after relocation
};
inline bool vm_run(const std::vector<fe>& program, uint64_t entry_offset, uint64_t max_steps, VmResult* out, std::string* err,
                   unsigned builtins = 0) {
    const uint64_t P = program.size();
    const uint64_t exec_base = 1 + P;
    const uint64_t SENT_FP = (1ULL << 62), SENT_PC = (1ULL << 62) + 1;
    // builtin segments live at sentinel bases until relocation; a cell value inside such a window is a pointer into the segment
    const uint64_t SEG_BASE[2] = {1ULL << 60, (1ULL << 60) + (1ULL << 50)}, SEG_SPAN = 1ULL << 24;
    const unsigned SEG_BIT[2] = {BUILTIN_OUTPUT, BUILTIN_RANGE_CHECK};
    std::vector<fe> seg_mem[2];
    std::vector<uint8_t> seg_known[2];
    auto seg_of = [&](uint64_t a) -> int {
        for (int s = 0; s < 2; ++s) if ((builtins & SEG_BIT[s]) && a >= SEG_BASE[s] && a < SEG_BASE[s] + SEG_SPAN) return s;
        return -1;
    };
    // Flat memory: a stray write far away must not allocate gigabytes.  The segments of a run are contiguous after
    // relocation, so every legitimate address is below program + 4 cells per step (+ slack); S252_CAIRO_MAX_ADDRESS raises the cap.
    uint64_t MAX_ADDRESS = std::min<uint64_t>(1ULL << 28, (uint64_t)program.size() + 4 * std::min<uint64_t>(max_steps, 1ULL << 26) + (1u << 16));
    if (const char* e = std::getenv("S252_CAIRO_MAX_ADDRESS")) { const unsigned long long v = std::strtoull(e, nullptr, 0); if (v) MAX_ADDRESS = v; }
    std::vector<fe> mem;          // index = address
    std::vector<uint8_t> known;
    auto ensure = [&](uint64_t a) { if (a >= mem.size()) { size_t n = std::max<size_t>(a + 1, mem.size() * 2); mem.resize(n, fe_zero()); known.resize(n, 0); } };
    auto set = [&](uint64_t a, const fe& v) -> bool {
        const int sg = seg_of(a);
        if (sg >= 0) {
            const uint64_t o = a - SEG_BASE[sg];
            if (o >= seg_mem[sg].size()) { seg_mem[sg].resize(o + 1, fe_zero()); seg_known[sg].resize(o + 1, 0); }
            if (seg_known[sg][o]) return H::eq(seg_mem[sg][o], v);
            seg_mem[sg][o] = v; seg_known[sg][o] = 1;
            return true;
        }
        if (a == 0 || a >= MAX_ADDRESS) return false;
        ensure(a);
        if (known[a]) return H::eq(mem[a], v);
        mem[a] = v; known[a] = 1;
        return true;
    };
    auto has = [&](uint64_t a) {
        const int sg = seg_of(a);
        if (sg >= 0) { const uint64_t o = a - SEG_BASE[sg]; return o < seg_mem[sg].size() && seg_known[sg][o] != 0; }
        return a < mem.size() && known[a];
    };
    auto get = [&](uint64_t a) -> const fe& {
        const int sg = seg_of(a);
        return sg >= 0 ? seg_mem[sg][a - SEG_BASE[sg]] : mem[a];
    };
    unsigned n_stack = 0;
    ensure(exec_base + 4);
    for (uint64_t i = 0; i < P; ++i) set(1 + i, program[i]);
    for (int sg = 0; sg < 2; ++sg) if (builtins & SEG_BIT[sg]) set(exec_base + n_stack++, H::from_u64(SEG_BASE[sg]));
    set(exec_base + n_stack, H::from_u64(SENT_FP));
    set(exec_base + n_stack + 1, H::from_u64(SENT_PC));
    uint64_t pc = 1 + entry_offset, ap = exec_base + n_stack + 2, fp = exec_base + n_stack + 2;
    uint64_t max_exec = exec_base + n_stack + 1;
    out->trace.clear();
    while (pc != SENT_PC) {
        if (out->trace.size() >= max_steps) { *err = "step limit reached"; return false; }
        if (!has(pc)) { *err = "InstructionNotFound"; return false; }
        Instr in;
        if (!decode(get(pc), &in, err)) return false;
        out->trace.push_back({pc, fp, ap});
        const uint64_t dst_addr = (in.bit(F_DST_FP) ? fp : ap) + (int64_t)in.soff[0];
        const uint64_t op0_addr = (in.bit(F_OP_0_FP) ? fp : ap) + (int64_t)in.soff[1];
        const unsigned size = in.size();
        // call writes its two operands before anything else is read
        if (in.opcode() == 1) {
            if (!set(dst_addr, H::from_u64(fp)) || !set(op0_addr, H::from_u64(pc + size))) { *err = "inconsistent call operands"; return false; }
        }
        uint64_t op1_addr;
        switch (in.op1_src()) {
            case 0:
                if (!has(op0_addr)) { *err = "op0 unknown for double dereference"; return false; }
                op1_addr = low64(get(op0_addr)) + (int64_t)in.soff[2];
                break;
            case 1: op1_addr = pc + (int64_t)in.soff[2]; break;
            case 2: op1_addr = fp + (int64_t)in.soff[2]; break;
            default: op1_addr = ap + (int64_t)in.soff[2]; break;
        }
        // operand deduction for assert_eq (Cairo whitepaper section 8.4)
        if (in.opcode() == 4 && has(dst_addr)) {
            const fe dst = get(dst_addr);
            if (in.res_logic() == 0 && !has(op1_addr)) set(op1_addr, dst);
            if (in.res_logic() == 1) {
                if (!has(op1_addr) && has(op0_addr)) set(op1_addr, H::sub(dst, get(op0_addr)));
                else if (!has(op0_addr) && has(op1_addr)) set(op0_addr, H::sub(dst, get(op1_addr)));
            }
            if (in.res_logic() == 2) {
                if (!has(op1_addr) && has(op0_addr) && !H::is_zero(get(op0_addr))) set(op1_addr, H::mul(dst, H::inv(get(op0_addr))));
                else if (!has(op0_addr) && has(op1_addr) && !H::is_zero(get(op1_addr))) set(op0_addr, H::mul(dst, H::inv(get(op1_addr))));
            }
        }
        fe res = fe_zero();
        bool res_known = false;
        if (in.pc_update() != 4 && has(op1_addr) && (in.res_logic() == 0 || has(op0_addr))) {
            res = in.res_logic() == 0 ? get(op1_addr) : in.res_logic() == 1 ? H::add(get(op0_addr), get(op1_addr)) : H::mul(get(op0_addr), get(op1_addr));
            res_known = true;
        }
        if (in.opcode() == 4) {
            if (!res_known) { *err = "assert_eq: res cannot be computed"; return false; }
            if (!set(dst_addr, res)) { *err = "assert_eq failed"; return false; }
        }
        if (!has(dst_addr) || !has(op0_addr) || !has(op1_addr)) {
            // the reference's trace builder reads all three cells of every step (execution_trace.rs:459-560)
            *err = "unknown operand cell";
            return false;
        }
        for (uint64_t a : {dst_addr, op0_addr, op1_addr}) if (a >= exec_base && a < MAX_ADDRESS) max_exec = std::max(max_exec, a);
        // register updates
        uint64_t npc, nap, nfp;
        switch (in.pc_update()) {
            case 0: npc = pc + size; break;
            case 1: if (!res_known) { *err = "jmp abs: res unknown"; return false; } npc = low64(res); break;
            case 2: if (!res_known) { *err = "jmp rel: res unknown"; return false; } npc = low64(H::add(H::from_u64(pc), res)); break;
            default: npc = H::is_zero(get(dst_addr)) ? pc + size : low64(H::add(H::from_u64(pc), get(op1_addr))); break;
        }
        switch (in.ap_update()) {
            case 0: nap = ap + (in.opcode() == 1 ? 2 : 0); break;
            case 1: if (!res_known) { *err = "ap += res: res unknown"; return false; } nap = low64(H::add(H::from_u64(ap), res)); break;
            default: nap = ap + 1; break;
        }
        if (in.opcode() == 1) nfp = ap + 2;
        else if (in.opcode() == 2) nfp = low64(get(dst_addr));
        else nfp = fp;
        pc = npc; ap = nap; fp = nfp;
    }
    // relocation: the builtin segments follow the execution segment in declaration order, then the two empty segments
    // [return_fp], [end] (both start where the last segment ends)
    const uint64_t exec_size = max_exec + 1 - exec_base;
    uint64_t next = exec_base + exec_size, seg_reloc[2] = {0, 0};
    for (int sg = 0; sg < 2; ++sg) {
        if (!(builtins & SEG_BIT[sg])) continue;
        seg_reloc[sg] = next;
        next += seg_mem[sg].size();
    }
    out->has_output = (builtins & BUILTIN_OUTPUT) != 0;
    out->has_rc = (builtins & BUILTIN_RANGE_CHECK) != 0;
    out->output_range[0] = seg_reloc[0]; out->output_range[1] = seg_reloc[0] + seg_mem[0].size();
    out->rc_range[0] = seg_reloc[1]; out->rc_range[1] = seg_reloc[1] + seg_mem[1].size();
    const fe reloc = H::from_u64(next);
    const fe sfp = H::from_u64(SENT_FP), spc = H::from_u64(SENT_PC);
    auto relocate = [&](const fe& v) -> fe {
        if (H::eq(v, sfp) || H::eq(v, spc)) return reloc;
        if (builtins) {
            const H::U256 c = H::to_u256(H::from_mont(v));
            if (c.w[1] == 0 && c.w[2] == 0 && c.w[3] == 0)
                for (int sg = 0; sg < 2; ++sg)      // one past the end is a valid pointer (the final builtin pointer main returns)
                    if ((builtins & SEG_BIT[sg]) && c.w[0] >= SEG_BASE[sg] && c.w[0] <= SEG_BASE[sg] + SEG_SPAN) return H::from_u64(seg_reloc[sg] + (c.w[0] - SEG_BASE[sg]));
        }
        return v;
    };
    out->memory.clear();
    for (uint64_t a = 1; a < mem.size(); ++a) {
        if (!known[a]) continue;
        out->memory.push_back({a, relocate(mem[a])});
    }
    for (int sg = 0; sg < 2; ++sg) {
        if (!(builtins & SEG_BIT[sg])) continue;
        for (uint64_t o = 0; o < seg_mem[sg].size(); ++o) {
            if (!seg_known[sg][o]) { *err = "a builtin segment has an unwritten cell"; return false; }
            if (sg == 1) {      // RangeCheckBuiltinRunner: every value is below 2^128
                const H::U256 c = H::to_u256(H::from_mont(seg_mem[sg][o]));
                if (c.w[2] || c.w[3]) { *err = "range-check builtin: value out of range"; return false; }
            }
            out->memory.push_back({seg_reloc[sg] + o, relocate(seg_mem[sg][o])});
        }
    }
    out->program_size = P;
    return true;
}

// ---------------------------------------------------------------------------------------------
// PublicInputs (src/cairo/air.rs:155-215)
struct PublicInputs {
    uint64_t pc_init = 0, ap_init = 0, fp_init = 0, pc_final = 0, ap_final = 0;
    bool has_rc = false;
    uint16_t range_check_min = 0, range_check_max = 0;
    bool has_rc_segment = false, has_output_segment = false;
    uint64_t rc_segment[2] = {0, 0}, output_segment[2] = {0, 0};
    std::vector<std::pair<uint64_t, fe>> public_memory;   // sorted by address (the reference keeps a HashMap)
    uint64_t num_steps = 0;
    const fe* pub_get(uint64_t a) const {
        for (auto& kv : public_memory) if (kv.first == a) return &kv.second;
        return nullptr;
    }
};

struct Table {
    std::vector<fe> t;   // row-major
    size_t n_cols = 0;
    size_t n_rows() const { return n_cols ? t.size() / n_cols : 0; }
    fe* row(size_t i) { return &t[i * n_cols]; }
    const fe* row(size_t i) const { return &t[i * n_cols]; }
};

inline bool parse_trace_le(const uint8_t* b, size_t len, std::vector<RegisterState>* out) {   // register_states.rs:47-78
    if (len % 24) return false;
    out->resize(len / 24);
    auto rd = [&](size_t o) { uint64_t v = 0; for (int k = 7; k >= 0; --k) v = (v << 8) | b[o + k]; return v; };
    for (size_t i = 0; i < out->size(); ++i) (*out)[i] = RegisterState{rd(24 * i + 16), rd(24 * i + 8), rd(24 * i)};
    return true;
}
inline bool parse_memory_le(const uint8_t* b, size_t len, Memory* out) {   // cairo_mem.rs:35-61
    if (len % 40) return false;
    out->data.reserve(len / 40 * 2);
    for (size_t i = 0; i < len / 40; ++i) {
        uint64_t a = 0;
        for (int k = 7; k >= 0; --k) a = (a << 8) | b[40 * i + k];
        out->data[a] = fe_from_le32(b + 40 * i + 8);
    }
    return true;
}

// PublicInputs::from_regs_and_mem (air.rs:183-214)
inline bool public_inputs_from_regs_and_mem(const std::vector<RegisterState>& regs, const Memory& mem, size_t program_size,
                                            const uint64_t* rc_segment, const uint64_t* output_segment, PublicInputs* pi,
                                            std::string* err) {
    if (regs.empty()) { *err = "empty register trace"; return false; }
    std::map<uint64_t, fe> pm;
    for (uint64_t i = 1; i <= program_size; ++i) {
        const fe* v = mem.get(i);
        if (!v) { *err = "program cell missing from memory"; return false; }
        pm[i] = *v;
    }
    if (output_segment)
        for (uint64_t a = output_segment[0]; a < output_segment[1]; ++a) {
            const fe* v = mem.get(a);
            if (!v) { *err = "output cell missing from memory"; return false; }
            pm[a] = *v;
        }
    pi->public_memory.assign(pm.begin(), pm.end());
    pi->pc_init = regs[0].pc; pi->ap_init = regs[0].ap; pi->fp_init = regs[0].fp;
    pi->pc_final = regs.back().pc; pi->ap_final = regs.back().ap;
    pi->has_rc = false;
    pi->has_rc_segment = rc_segment != nullptr;
    pi->has_output_segment = output_segment != nullptr;
    if (rc_segment) { pi->rc_segment[0] = rc_segment[0]; pi->rc_segment[1] = rc_segment[1]; }
    if (output_segment) { pi->output_segment[0] = output_segment[0]; pi->output_segment[1] = output_segment[1]; }
    pi->num_steps = regs.size();
    return true;
}

// build_cairo_execution_trace (execution_trace.rs:261-356)
inline bool build_cairo_execution_trace(const std::vector<RegisterState>& regs, const Memory& mem, const PublicInputs& pi,
                                        Table* out, std::string* err) {
    const size_t n = regs.size();
    const bool rc_builtin = pi.has_rc_segment;
    const size_t cols = MAIN_COLS + (rc_builtin ? RC_BUILTIN_COLS : 0);
    out->n_cols = cols;
    {
        size_t p2 = 1;
        while (p2 < n + (n >> 3) + 64) p2 <<= 1;     // room for the hole / dummy / padding rows build_main_trace appends
        out->t.reserve(p2 * cols);
    }
    out->t.assign(n * cols, fe_zero());
    const fe one = H::one();
    std::vector<std::vector<size_t>> jnz_parts(64);   // rows whose res is dst^-1: inverted together after the loop
    std::vector<std::string> errs(64);
    parallel_for(n, [&](size_t lo_, size_t hi_, unsigned worker) {
    std::vector<size_t>& jnz_rows = jnz_parts[worker];
    std::string* err = &errs[worker];
    auto fail = [&](const char* m) { *err = m; };
    for (size_t i = lo_; i < hi_; ++i) {
        const RegisterState& s = regs[i];
        const fe* instp = mem.get(s.pc);
        if (!instp) { fail("InstructionNotFound"); return; }
        Instr in;
        if (!decode(*instp, &in, err)) return;
        fe* r = out->row(i);
        for (unsigned b = 0; b < 15; ++b) r[b] = in.bit(b) ? one : fe_zero();
        auto addr_of = [&](uint64_t base, int32_t off, uint64_t* a) -> bool {   // checked_add_signed().unwrap()
            if (off < 0 && base < (uint64_t)(-(int64_t)off)) return false;
            *a = base + (int64_t)off;
            return true;
        };
        uint64_t dst_addr, op0_addr, op1_addr;
        if (!addr_of(in.bit(F_DST_FP) ? s.fp : s.ap, in.soff[0], &dst_addr) || !addr_of(in.bit(F_OP_0_FP) ? s.fp : s.ap, in.soff[1], &op0_addr)) {
            fail("address underflow"); return;
        }
        const fe *dstp = mem.get(dst_addr), *op0p = mem.get(op0_addr);
        if (!dstp || !op0p) { fail("operand cell missing from memory"); return; }
        fe dst = *dstp, op0 = *op0p;
        uint64_t op1_base;
        switch (in.op1_src()) {
            case 0: op1_base = low64(op0); break;
            case 1: op1_base = s.pc; break;
            case 2: op1_base = s.fp; break;
            default: op1_base = s.ap; break;
        }
        if (!addr_of(op1_base, in.soff[2], &op1_addr)) { fail("address underflow"); return; }
        const fe* op1p = mem.get(op1_addr);
        if (!op1p) { fail("operand cell missing from memory"); return; }
        const fe op1 = *op1p;
        // compute_res (execution_trace.rs:381-440)
        fe res;
        if (in.pc_update() == 4) {
            if (!(in.res_logic() == 0 && in.opcode() == 0)) { fail("Undefined Behavior"); return; }
            res = dst;                                   // dst == 0 ? dst : dst.inv()  (the inversion is batched below)
            if (!H::is_zero(dst)) jnz_rows.push_back(i);
        } else {
            res = in.res_logic() == 0 ? op1 : in.res_logic() == 1 ? H::add(op0, op1) : H::mul(op0, op1);
        }
        // update_values (execution_trace.rs:565-585)
        if (in.opcode() == 1) { op0 = H::from_u64(s.pc + in.size()); dst = H::from_u64(s.fp); }
        else if (in.opcode() == 4) res = dst;
        r[FRAME_RES] = res;
        r[FRAME_AP] = H::from_u64(s.ap); r[FRAME_FP] = H::from_u64(s.fp); r[FRAME_PC] = H::from_u64(s.pc);
        r[FRAME_DST_ADDR] = H::from_u64(dst_addr); r[FRAME_OP0_ADDR] = H::from_u64(op0_addr); r[FRAME_OP1_ADDR] = H::from_u64(op1_addr);
        r[FRAME_INST] = *instp; r[FRAME_DST] = dst; r[FRAME_OP0] = op0; r[FRAME_OP1] = op1;
        for (int k = 0; k < 3; ++k) r[OFF_DST + k] = H::from_u64(in.off[k]);   // to_unbiased_representation
        const fe t0 = in.bit(F_PC_JNZ) ? dst : fe_zero();
        r[FRAME_T0] = t0;
        r[FRAME_T1] = H::mul(t0, res);
        r[FRAME_MUL] = H::mul(op0, op1);
        r[FRAME_SELECTOR] = (i + 1 == n) ? fe_zero() : one;
    }
    });
    for (auto& e : errs) if (!e.empty()) { *err = e; return false; }
    std::vector<size_t> jnz_rows;
    for (auto& part : jnz_parts) jnz_rows.insert(jnz_rows.end(), part.begin(), part.end());
    if (!jnz_rows.empty()) {   // one field inversion for all jnz rows (Montgomery's trick); t1 = t0 * res follows
        std::vector<fe> pre(jnz_rows.size());
        fe acc = one;
        for (size_t k = 0; k < jnz_rows.size(); ++k) { pre[k] = acc; acc = H::mul(acc, out->row(jnz_rows[k])[FRAME_RES]); }
        fe inv = H::inv(acc);
        for (size_t k = jnz_rows.size(); k-- > 0;) {
            fe* r = out->row(jnz_rows[k]);
            const fe d = r[FRAME_RES];
            r[FRAME_RES] = H::mul(inv, pre[k]);
            inv = H::mul(inv, d);
            r[FRAME_T1] = H::mul(r[FRAME_T0], r[FRAME_RES]);
        }
    }
    if (rc_builtin) {   // add_rc_builtin_columns (execution_trace.rs:358-379)
        size_t k = 0;
        for (uint64_t a = pi.rc_segment[0]; a < pi.rc_segment[1]; ++a, ++k) {
            const fe* v = mem.get(a);
            if (!v) { *err = "range-check cell missing from memory"; return false; }
            if (k >= n) break;   // column.resize(trace_len) truncates
            const H::U256 c = H::to_u256(H::from_mont(*v));
            fe* r = out->row(k);
            for (int j = 0; j < 8; ++j) r[MAIN_COLS + j] = H::from_u64((c.w[j / 4] >> (16 * (j % 4))) & 0xffff);
            r[MAIN_COLS + 8] = *v;
        }
    }
    return true;
}

inline void pad_rows(Table* t, size_t count, const std::vector<fe>& row) {
    for (size_t k = 0; k < count; ++k) t->t.insert(t->t.end(), row.begin(), row.end());
}

// build_main_trace (execution_trace.rs:57-87)
inline bool build_main_trace(const std::vector<RegisterState>& regs, const Memory& mem, PublicInputs* pi, Table* out, std::string* err) {
    if (!build_cairo_execution_trace(regs, mem, *pi, out, err)) return false;
    Table& T = *out;
    const size_t cols = T.n_cols;
    const unsigned ADDR_COLUMNS[4] = {FRAME_PC, FRAME_DST_ADDR, FRAME_OP0_ADDR, FRAME_OP1_ADDR};
    const unsigned MEMORY_COLUMNS[8] = {FRAME_PC, FRAME_DST_ADDR, FRAME_OP0_ADDR, FRAME_OP1_ADDR, FRAME_INST, FRAME_DST, FRAME_OP0, FRAME_OP1};
    // sorted addresses of the execution rows (addresses are machine words)
    std::vector<uint64_t> addrs;
    addrs.reserve(T.n_rows() * 4);
    for (size_t i = 0; i < T.n_rows(); ++i)
        for (unsigned c : ADDR_COLUMNS) addrs.push_back(low64(T.row(i)[c]));
    std::sort(addrs.begin(), addrs.end());
    // get_rc_holes (execution_trace.rs:136-174) over the three offset columns
    std::vector<uint16_t> offs;
    offs.reserve(T.n_rows() * 3);
    for (size_t i = 0; i < T.n_rows(); ++i)
        for (unsigned c = OFF_DST; c <= OFF_OP1; ++c) offs.push_back((uint16_t)low64(T.row(i)[c]));
    std::sort(offs.begin(), offs.end());
    std::vector<uint64_t> holes;
    for (size_t i = 0; i + 1 < offs.size(); ++i)
        for (uint32_t v = (uint32_t)offs[i] + 1; v < offs[i + 1]; ++v) holes.push_back(v);
    const size_t pad3 = ((holes.size() + 2) / 3) * 3 - holes.size();
    for (size_t i = 0; i < pad3; ++i) holes.push_back(offs.back());
    pi->has_rc = true;
    pi->range_check_min = offs.front();
    pi->range_check_max = offs.back();
    // fill_rc_holes (execution_trace.rs:176-186)
    for (size_t i = 0; i < holes.size(); i += 3) {
        std::vector<fe> row(cols, fe_zero());
        for (int k = 0; k < 3; ++k) row[OFF_DST + k] = H::from_u64(holes[i + k]);
        pad_rows(&T, 1, row);
    }
    // get_memory_holes (execution_trace.rs:196-222), fill_memory_holes (:227-259)
    const uint64_t codelen = pi->public_memory.size();
    std::vector<uint64_t> mholes;
    uint64_t prev = addrs[0];
    // a caller-supplied memory file may contain a gap of 2^40 addresses: holes beyond a few times the trace length
    // cannot belong to a real execution (every hole costs a third of a padding row) -- refuse instead of exhausting RAM
    const size_t max_holes = 16 * (size_t)T.n_rows() + (1u << 16);
    for (uint64_t a : addrs) {
        const uint64_t diff = a - prev;
        if (diff != 1 && diff != 0 && a > codelen) {
            if (diff > max_holes || mholes.size() + diff > max_holes) { *err = "memory holes exceed 16x the trace length"; return false; }
            for (uint64_t h = prev + 1; h < a; ++h)
                if (h > codelen) mholes.push_back(h);
        }
        prev = a;
    }
    if (!mholes.empty()) {
        const std::vector<fe> last(T.row(T.n_rows() - 1), T.row(T.n_rows() - 1) + cols);
        size_t k = 0;
        while (k < mholes.size()) {
            std::vector<fe> row = last;
            for (unsigned c : ADDR_COLUMNS)
                if (k < mholes.size()) row[c] = H::from_u64(mholes[k++]);
            pad_rows(&T, 1, row);
        }
    }
    // add_pub_memory_dummy_accesses (execution_trace.rs:91-96)
    {
        std::vector<fe> row(T.row(T.n_rows() - 1), T.row(T.n_rows() - 1) + cols);
        for (unsigned c : MEMORY_COLUMNS) row[c] = fe_zero();
        pad_rows(&T, (pi->public_memory.size() >> 2) + 1, row);
    }
    // pad_with_last_row up to the next power of two
    {
        size_t n = T.n_rows(), p2 = 1;
        while (p2 < n) p2 <<= 1;
        const std::vector<fe> row(T.row(n - 1), T.row(n - 1) + cols);
        pad_rows(&T, p2 - n, row);
    }
    return true;
}

}  // namespace cairo
}  // namespace s252
