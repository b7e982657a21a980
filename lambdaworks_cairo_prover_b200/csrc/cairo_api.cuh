// cairo_api.cuh -- C ABI of include/stark252_cairo.h (included at the end of runtime.cu: it uses the
// context, arena and transform helpers defined there).
#pragma once
#include "../../include/stark252_cairo.h"
#include "cairo_host.hpp"

namespace CA = s252::cairo;

static thread_local std::string g_cairo_err;
#define CAIRO_FAIL(code, msg) do { g_cairo_err = (msg); return (code); } while (0)

struct s252_cairo_run {
    CA::VmResult r;
};
struct s252_cairo_trace {
    std::vector<s252_fe> table;   // row-major, LW
    size_t n_rows = 0, n_cols = 0;
    CA::PublicInputs pi;
};

extern "C" const char* s252_cairo_last_error(void) { return g_cairo_err.c_str(); }

extern "C" int s252_cairo_vm_run(const uint8_t* program_be, size_t n_words, uint64_t entry_offset, uint64_t max_steps,
                                 s252_cairo_run** out) {
    if (!program_be || !n_words || !out || entry_offset >= n_words) CAIRO_FAIL(S252_ERR_INVALID, "bad program");
    std::vector<fe> prog(n_words);
    for (size_t i = 0; i < n_words; ++i) prog[i] = H::from_bytes_be(program_be + 32 * i);
    s252_cairo_run* run = new s252_cairo_run();
    std::string err;
    if (!CA::vm_run(prog, entry_offset, max_steps ? max_steps : ~0ULL, &run->r, &err)) {
        delete run;
        CAIRO_FAIL(S252_ERR_INVALID, "cairo vm: " + err);
    }
    *out = run;
    return S252_OK;
}
extern "C" void s252_cairo_run_destroy(s252_cairo_run* run) { delete run; }
extern "C" size_t s252_cairo_run_steps(const s252_cairo_run* run) { return run->r.trace.size(); }
extern "C" size_t s252_cairo_run_trace_len(const s252_cairo_run* run) { return run->r.trace.size() * 24; }
extern "C" size_t s252_cairo_run_memory_len(const s252_cairo_run* run) { return run->r.memory.size() * 40; }
static inline void put_u64_le(uint8_t* o, uint64_t v) { for (int k = 0; k < 8; ++k) o[k] = (uint8_t)(v >> (8 * k)); }
static inline void put_u64_be(uint8_t* o, uint64_t v) { for (int k = 0; k < 8; ++k) o[k] = (uint8_t)(v >> (56 - 8 * k)); }
extern "C" void s252_cairo_run_trace_bytes(const s252_cairo_run* run, uint8_t* out) {
    for (size_t i = 0; i < run->r.trace.size(); ++i) {
        put_u64_le(out + 24 * i, run->r.trace[i].ap);
        put_u64_le(out + 24 * i + 8, run->r.trace[i].fp);
        put_u64_le(out + 24 * i + 16, run->r.trace[i].pc);
    }
}
extern "C" void s252_cairo_run_memory_bytes(const s252_cairo_run* run, uint8_t* out) {
    for (size_t i = 0; i < run->r.memory.size(); ++i) {
        put_u64_le(out + 40 * i, run->r.memory[i].first);
        CA::fe_to_le32(run->r.memory[i].second, out + 40 * i + 8);
    }
}

static int cairo_build_common(bool full, const uint8_t* trace_le, size_t trace_len, const uint8_t* memory_le, size_t memory_len,
                              size_t program_size, const uint64_t* rc_range, const uint64_t* output_range, s252_cairo_trace** out) {
    if (!trace_le || !memory_le || !out) CAIRO_FAIL(S252_ERR_INVALID, "null argument");
    std::vector<CA::RegisterState> regs;
    CA::Memory mem;
    if (!CA::parse_trace_le(trace_le, trace_len, &regs)) CAIRO_FAIL(S252_ERR_INVALID, "IncorrectNumberOfBytes (register trace)");
    if (!CA::parse_memory_le(memory_le, memory_len, &mem)) CAIRO_FAIL(S252_ERR_INVALID, "IncorrectNumberOfBytes (memory)");
    s252_cairo_trace* t = new s252_cairo_trace();
    std::string err;
    CA::Table tab;
    bool ok = CA::public_inputs_from_regs_and_mem(regs, mem, program_size, rc_range, output_range, &t->pi, &err);
    if (ok) ok = full ? CA::build_main_trace(regs, mem, &t->pi, &tab, &err) : CA::build_cairo_execution_trace(regs, mem, t->pi, &tab, &err);
    if (!ok) {
        delete t;
        CAIRO_FAIL(S252_ERR_INVALID, "build_main_trace: " + err);
    }
    t->n_cols = tab.n_cols;
    t->n_rows = tab.n_rows();
    t->table.resize(tab.t.size());
    for (size_t i = 0; i < tab.t.size(); ++i) H::to_lw(tab.t[i], t->table[i].limbs);
    *out = t;
    return S252_OK;
}
extern "C" int s252_cairo_build_main_trace(const uint8_t* trace_le, size_t trace_len, const uint8_t* memory_le, size_t memory_len,
                                           size_t program_size, const uint64_t* rc_range, const uint64_t* output_range,
                                           s252_cairo_trace** out) {
    return cairo_build_common(true, trace_le, trace_len, memory_le, memory_len, program_size, rc_range, output_range, out);
}
extern "C" int s252_cairo_build_execution_trace(const uint8_t* trace_le, size_t trace_len, const uint8_t* memory_le,
                                                size_t memory_len, size_t program_size, const uint64_t* rc_range,
                                                const uint64_t* output_range, s252_cairo_trace** out) {
    return cairo_build_common(false, trace_le, trace_len, memory_le, memory_len, program_size, rc_range, output_range, out);
}
extern "C" void s252_cairo_trace_destroy(s252_cairo_trace* t) { delete t; }
extern "C" size_t s252_cairo_trace_n_rows(const s252_cairo_trace* t) { return t->n_rows; }
extern "C" size_t s252_cairo_trace_n_cols(const s252_cairo_trace* t) { return t->n_cols; }
extern "C" const s252_fe* s252_cairo_trace_table(const s252_cairo_trace* t) { return t->table.data(); }
extern "C" void s252_cairo_trace_public_inputs(const s252_cairo_trace* t, s252_cairo_public_inputs* o) {
    const CA::PublicInputs& p = t->pi;
    std::memset(o, 0, sizeof *o);
    o->pc_init = p.pc_init; o->ap_init = p.ap_init; o->fp_init = p.fp_init; o->pc_final = p.pc_final; o->ap_final = p.ap_final;
    o->num_steps = p.num_steps;
    o->n_public_memory = p.public_memory.size();
    o->rc_segment[0] = p.rc_segment[0]; o->rc_segment[1] = p.rc_segment[1];
    o->output_segment[0] = p.output_segment[0]; o->output_segment[1] = p.output_segment[1];
    o->range_check_min = p.range_check_min; o->range_check_max = p.range_check_max;
    o->has_range_check_bounds = p.has_rc; o->has_rc_segment = p.has_rc_segment; o->has_output_segment = p.has_output_segment;
}
extern "C" void s252_cairo_trace_public_memory(const s252_cairo_trace* t, uint64_t* addrs, s252_fe* values) {
    for (size_t i = 0; i < t->pi.public_memory.size(); ++i) {
        addrs[i] = t->pi.public_memory[i].first;
        H::to_lw(t->pi.public_memory[i].second, values[i].limbs);
    }
}
static void serialize_public_inputs(const CA::PublicInputs& p, std::vector<uint8_t>* out) {   // air.rs:217-276
    auto u64be = [&](uint64_t v) { uint8_t b[8]; put_u64_be(b, v); out->insert(out->end(), b, b + 8); };
    auto felt = [&](const fe& v) { uint8_t b[32]; H::to_bytes_be(v, b); out->insert(out->end(), b, b + 32); };
    u64be(32);
    for (uint64_t v : {p.pc_init, p.ap_init, p.fp_init, p.pc_final, p.ap_final}) felt(H::from_u64(v));
    for (uint16_t v : {p.range_check_min, p.range_check_max}) {
        if (p.has_rc) { out->push_back(1); out->push_back((uint8_t)(v >> 8)); out->push_back((uint8_t)v); }
        else out->push_back(0);
    }
    std::vector<uint8_t> seg;
    size_t nseg = 0;
    auto put_seg = [&](uint8_t kind, const uint64_t r[2]) {
        seg.push_back(kind);
        uint8_t b[8];
        put_u64_be(b, r[0]); seg.insert(seg.end(), b, b + 8);
        put_u64_be(b, r[1]); seg.insert(seg.end(), b, b + 8);
        ++nseg;
    };
    if (p.has_rc_segment) put_seg(0, p.rc_segment);
    if (p.has_output_segment) put_seg(1, p.output_segment);
    u64be(nseg);
    out->insert(out->end(), seg.begin(), seg.end());
    u64be(p.public_memory.size());
    for (auto& kv : p.public_memory) { felt(H::from_u64(kv.first)); felt(kv.second); }
    u64be(p.num_steps);
}
extern "C" size_t s252_cairo_trace_serialize_public_inputs(const s252_cairo_trace* t, uint8_t* out) {
    std::vector<uint8_t> b;
    serialize_public_inputs(t->pi, &b);
    if (out) std::memcpy(out, b.data(), b.size());
    return b.size();
}
