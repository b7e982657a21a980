// cairo_api.cuh -- C ABI of include/stark252_cairo.h (included at the end of runtime.cu: it uses the
// context, arena and transform helpers defined there).
#pragma once
#include "../../include/stark252_cairo.h"
#include "cairo_host.hpp"
#include <chrono>

namespace CA = s252::cairo;

static thread_local std::string g_cairo_err;
#define CAIRO_FAIL(code, msg) do { g_cairo_err = (msg); return (code); } while (0)

struct s252_cairo_run {
    CA::VmResult r;
};
struct s252_cairo_trace {
    std::vector<s252_fe> table;   // row-major, LW (TraceTable.table)
    std::vector<s252_fe> cols;    // the same table column-major: what round 1 uploads, one group of columns at a time
    size_t n_rows = 0, n_cols = 0;
    CA::PublicInputs pi;
    mutable bool pinned = false;  // `cols` pages registered with the CUDA driver (true DMA for the upload of round 1)
    void make_columns() {
        cols.resize(table.size());
        CA::parallel_for(n_rows, [&](size_t lo, size_t hi, unsigned) {
            for (size_t i0 = lo; i0 < hi; i0 += 64)          // 64-row blocks: the strided writes stay within a few pages per column
                for (size_t j = 0; j < n_cols; ++j)
                    for (size_t i = i0; i < std::min(hi, i0 + 64); ++i) cols[j * n_rows + i] = table[i * n_cols + j];
        });
    }
    ~s252_cairo_trace() { if (pinned) cudaHostUnregister((void*)cols.data()); }
};

extern "C" const char* s252_cairo_last_error(void) { return g_cairo_err.c_str(); }

static int cairo_vm_run_impl(const uint8_t* program_be, size_t n_words, uint64_t entry_offset, uint64_t max_steps, unsigned builtins,
                             s252_cairo_run** out);
extern "C" int s252_cairo_vm_run(const uint8_t* program_be, size_t n_words, uint64_t entry_offset, uint64_t max_steps,
                                 s252_cairo_run** out) {
    return cairo_vm_run_impl(program_be, n_words, entry_offset, max_steps, 0, out);
}
extern "C" int s252_cairo_vm_run_builtins(const uint8_t* program_be, size_t n_words, uint64_t entry_offset, uint64_t max_steps,
                                          unsigned builtins, s252_cairo_run** out) {
    if (builtins & ~(CA::BUILTIN_OUTPUT | CA::BUILTIN_RANGE_CHECK)) CAIRO_FAIL(S252_ERR_INVALID, "unsupported builtin");
    return cairo_vm_run_impl(program_be, n_words, entry_offset, max_steps, builtins, out);
}
extern "C" int s252_cairo_run_segment(const s252_cairo_run* run, int which, uint64_t range[2]) {
    if (!run || !range) return 0;
    const bool has = which == 0 ? run->r.has_rc : run->r.has_output;
    const uint64_t* r = which == 0 ? run->r.rc_range : run->r.output_range;
    range[0] = r[0]; range[1] = r[1];
    return has ? 1 : 0;
}
static int cairo_vm_run_impl(const uint8_t* program_be, size_t n_words, uint64_t entry_offset, uint64_t max_steps, unsigned builtins,
                             s252_cairo_run** out) {
    if (!program_be || !n_words || !out || entry_offset >= n_words) CAIRO_FAIL(S252_ERR_INVALID, "bad program");
    s252_cairo_run* run = nullptr;
    try {
        std::vector<fe> prog(n_words);
        for (size_t i = 0; i < n_words; ++i) prog[i] = H::from_bytes_be(program_be + 32 * i);
        run = new s252_cairo_run();
        std::string err;
        if (!CA::vm_run(prog, entry_offset, max_steps ? max_steps : ~0ULL, &run->r, &err, builtins)) {
            delete run;
            CAIRO_FAIL(S252_ERR_INVALID, "cairo vm: " + err);
        }
    } catch (const std::exception& e) {           // nothing may unwind across the C boundary
        delete run;
        CAIRO_FAIL(S252_ERR_INVALID, std::string("cairo vm: ") + e.what());
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
    const bool timing = std::getenv("S252_TIMING") != nullptr;
    auto t_prev = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!timing) return;
        const auto t_now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[s252 front-end] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(t_now - t_prev).count());
        t_prev = t_now;
    };
    std::vector<CA::RegisterState> regs;
    CA::Memory mem;
    if (!CA::parse_trace_le(trace_le, trace_len, &regs)) CAIRO_FAIL(S252_ERR_INVALID, "IncorrectNumberOfBytes (register trace)");
    if (!CA::parse_memory_le(memory_le, memory_len, &mem)) CAIRO_FAIL(S252_ERR_INVALID, "IncorrectNumberOfBytes (memory)");
    lap("parse trace + memory files");
    s252_cairo_trace* t = nullptr;
    try {
        t = new s252_cairo_trace();
        std::string err;
        CA::Table tab;
        bool ok = CA::public_inputs_from_regs_and_mem(regs, mem, program_size, rc_range, output_range, &t->pi, &err);
        if (ok) ok = full ? CA::build_main_trace(regs, mem, &t->pi, &tab, &err) : CA::build_cairo_execution_trace(regs, mem, t->pi, &tab, &err);
        lap("build_main_trace");
        if (!ok) {
            delete t;
            CAIRO_FAIL(S252_ERR_INVALID, "build_main_trace: " + err);
        }
        t->n_cols = tab.n_cols;
        t->n_rows = tab.n_rows();
        t->table.resize(tab.t.size());
        CA::parallel_for(tab.t.size(), [&](size_t lo, size_t hi, unsigned) {
            for (size_t i = lo; i < hi; ++i) H::to_lw(tab.t[i], t->table[i].limbs);
        });
        lap("table in the LW layout");
        t->make_columns();
        lap("column-major copy");
    } catch (const std::exception& e) {
        delete t;
        CAIRO_FAIL(S252_ERR_INVALID, std::string("build_main_trace: ") + e.what());
    }
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
extern "C" int s252_cairo_trace_pin(const s252_cairo_trace* t) {
    if (!t) return S252_ERR_INVALID;
    if (!t->pinned && !t->cols.empty()) {
        if (cudaHostRegister((void*)t->cols.data(), t->cols.size() * sizeof(s252_fe), cudaHostRegisterDefault) != cudaSuccess) {
            cudaGetLastError();
            CAIRO_FAIL(S252_ERR_CUDA, "cudaHostRegister failed");
        }
        t->pinned = true;
    }
    return S252_OK;
}
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

extern "C" int s252_cairo_trace_from_table(const s252_fe* table, size_t n_rows, size_t n_cols, const s252_cairo_public_inputs* pub,
                                           const uint64_t* pub_addrs, const s252_fe* pub_values, s252_cairo_trace** out) {
    if (!table || !pub || !out || (pub->n_public_memory && (!pub_addrs || !pub_values))) CAIRO_FAIL(S252_ERR_INVALID, "null argument");
    if (n_cols != CA::MAIN_COLS && n_cols != CA::MAIN_COLS + CA::RC_BUILTIN_COLS) CAIRO_FAIL(S252_ERR_INVALID, "a Cairo main trace has 34 or 43 columns");
    if (n_rows == 0 || !is_pow2(n_rows) || n_rows > ((size_t)1 << 40) / n_cols) CAIRO_FAIL(S252_ERR_INVALID, "the trace length must be a power of two (and n_rows * n_cols must not overflow)");
    s252_cairo_trace* t = nullptr;
    try {                                                                  // nothing may unwind across the C boundary
        t = new s252_cairo_trace();
        t->table.assign(table, table + n_rows * n_cols);
        t->n_rows = n_rows; t->n_cols = n_cols;
        CA::PublicInputs& p = t->pi;
        p.pc_init = pub->pc_init; p.ap_init = pub->ap_init; p.fp_init = pub->fp_init; p.pc_final = pub->pc_final; p.ap_final = pub->ap_final;
        p.num_steps = pub->num_steps;
        p.has_rc = pub->has_range_check_bounds; p.range_check_min = pub->range_check_min; p.range_check_max = pub->range_check_max;
        p.has_rc_segment = pub->has_rc_segment; p.has_output_segment = pub->has_output_segment;
        p.rc_segment[0] = pub->rc_segment[0]; p.rc_segment[1] = pub->rc_segment[1];
        p.output_segment[0] = pub->output_segment[0]; p.output_segment[1] = pub->output_segment[1];
        for (size_t i = 0; i < pub->n_public_memory; ++i) p.public_memory.push_back({pub_addrs[i], H::from_lw(pub_values[i].limbs)});
        std::sort(p.public_memory.begin(), p.public_memory.end(), [](const std::pair<uint64_t, fe>& a, const std::pair<uint64_t, fe>& b) { return a.first < b.first; });
        t->make_columns();
    } catch (const std::exception& e) {
        delete t;
        CAIRO_FAIL(S252_ERR_INVALID, std::string("s252_cairo_trace_from_table: ") + e.what());
    }
    *out = t;
    return S252_OK;
}

// --------------------------------------------------------------------------------------------
// GPU prover

// wall-clock stage timer of the last s252_cairo_prove on this thread (diagnostics; s252_cairo_last_prove_stages)
#include <chrono>
static thread_local std::string g_cairo_stages, g_cairo_round1;
struct StageTimer {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    std::string json = "{";
    void mark(const char* name) {
        const auto t1 = std::chrono::steady_clock::now();
        char b[96];
        std::snprintf(b, sizeof b, "%s\"%s\": %.3f", json.size() > 1 ? ", " : "", name, std::chrono::duration<double, std::milli>(t1 - t0).count());
        json += b;
        t0 = t1;
    }
};
extern "C" const char* s252_cairo_last_prove_stages(void) { return g_cairo_stages.c_str(); }


// dom[i] = h w^i and T[i] = 1/(dom[i] - 1) over the LDE coset; both cached in the context
static int get_coset_tables(s252_ctx* ctx, size_t M, uint64_t coset_offset, const fe** dom, const fe** T) {
    fe w;
    if (!H::primitive_root(ilog2(M), &w)) FAIL(ctx, S252_ERR_INVALID, "no root of unity of order %zu", M);
    const fe h = H::from_u64(coset_offset);
    TRY(get_power_table(ctx, M, w, h, dom));
    fe* t; bool fresh;
    TRY(table_alloc(ctx, "cairoT:" + std::to_string(M) + ":" + fe_key(h), M, &t, &fresh));
    if (fresh) {
        TRY(invert_shifted(ctx, *dom, M, H::one(), t));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
    }
    *T = t;
    return S252_OK;
}

// Stable LSD radix sort over the low `bits` bits (kernels in cairo.cuh).  Ping-pongs between the two buffers;
// returns through *result which one holds the sorted keys (vals likewise).
template <typename K>
static int radix_sort(s252_ctx* ctx, K* keys, K* keys2, unsigned* vals, unsigned* vals2, size_t n, unsigned bits, unsigned* hist,
                      int* result) {
    const unsigned units = (unsigned)((n + s252::RS_ITEMS - 1) / s252::RS_ITEMS);
    const unsigned blocks = (units + s252::RS_WARPS - 1) / s252::RS_WARPS;
    int cur = 0;
    for (unsigned shift = 0; shift < bits; shift += 8) {
        K* kin = cur ? keys2 : keys; K* kout = cur ? keys : keys2;
        unsigned* vin = cur ? vals2 : vals; unsigned* vout = cur ? vals : vals2;
        prof_begin(ctx, "radix_sort_pass");
        prof_work(ctx, (double)n * (2 * sizeof(K) + (vals ? 8 : 0) + sizeof(K)), 0, 0);
        s252::rs_histogram<K><<<blocks, s252::RS_WARPS * 32, 0, ctx->stream>>>(kin, (unsigned)n, shift, hist, units);
        s252::rs_scan_rows<<<256, 1024, 0, ctx->stream>>>(hist, units, hist + 256 * units);
        if (vals) s252::rs_scatter<K, true><<<blocks, s252::RS_WARPS * 32, 0, ctx->stream>>>(kin, vin, (unsigned)n, shift, hist, units, hist + 256 * units, kout, vout);
        else s252::rs_scatter<K, false><<<blocks, s252::RS_WARPS * 32, 0, ctx->stream>>>(kin, nullptr, (unsigned)n, shift, hist, units, hist + 256 * units, kout, nullptr);
        ctx->launches += 2;
        LAUNCH_CHECK(ctx);
        cur ^= 1;
    }
    *result = cur;
    return S252_OK;
}

// build_auxiliary_trace (air.rs:660-729) into aux[18][N] (column-major, internal format)
static int cairo_build_aux(s252_ctx* ctx, const fe* main_cols, unsigned col0, size_t N, const CA::PublicInputs& pi, const fe rap[3], fe* aux) {
    const size_t L = 4 * N, R = 3 * N, np = pi.public_memory.size();
    if (np > L) FAIL(ctx, S252_ERR_INVALID, "public memory (%zu words) does not fit the trace", np);
    if (L >= (1ull << 31)) FAIL(ctx, S252_ERR_INVALID, "trace too long for the auxiliary-trace sort");
    std::vector<unsigned long long> paddr(np);
    std::vector<fe> paddr_fe(np), pval(np);
    for (size_t i = 0; i < np; ++i) { paddr[i] = pi.public_memory[i].first; paddr_fe[i] = H::from_u64(paddr[i]); pval[i] = pi.public_memory[i].second; }
    Tmp<unsigned long long> d_paddr(ctx), keys(ctx), keys2(ctx);
    Tmp<fe> d_paddr_fe(ctx), d_pval(ctx), num(ctx), den(ctx), rnum(ctx), rden(ctx);
    Tmp<unsigned> idx(ctx), idx2(ctx);
    Tmp<unsigned short> okeys(ctx), okeys2(ctx);
    Tmp<unsigned> hist(ctx);
    Tmp<unsigned long long> key_or(ctx);
    TRY(dalloc(ctx, &d_paddr.p, np + 1)); TRY(dalloc(ctx, &d_paddr_fe.p, np + 1)); TRY(dalloc(ctx, &d_pval.p, np + 1));
    if (np) {
        CU(ctx, cudaMemcpyAsync(d_paddr.p, paddr.data(), np * 8, cudaMemcpyHostToDevice, ctx->stream));
        CU(ctx, cudaMemcpyAsync(d_paddr_fe.p, paddr_fe.data(), np * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream));
        CU(ctx, cudaMemcpyAsync(d_pval.p, pval.data(), np * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream));
    }
    TRY(dalloc(ctx, &keys.p, L)); TRY(dalloc(ctx, &keys2.p, L)); TRY(dalloc(ctx, &idx.p, L)); TRY(dalloc(ctx, &idx2.p, L));
    TRY(dalloc(ctx, &okeys.p, R)); TRY(dalloc(ctx, &okeys2.p, R));
    TRY(dalloc(ctx, &num.p, L)); TRY(dalloc(ctx, &den.p, L)); TRY(dalloc(ctx, &rnum.p, R)); TRY(dalloc(ctx, &rden.p, R));
    TRY(dalloc(ctx, &hist.p, 256 * ((L + s252::RS_ITEMS - 1) / s252::RS_ITEMS) + 256));   // (digit, warp) counts + 256 digit totals
    TRY(dalloc(ctx, &key_or.p, 1));
    CU(ctx, cudaMemsetAsync(key_or.p, 0, 8, ctx->stream));
    s252::CairoAux P{};
    P.main = main_cols; P.col0 = col0; P.n = N; P.pub_addr = d_paddr.p; P.pub_addr_fe = d_paddr_fe.p; P.pub_val = d_pval.p; P.n_pub = (unsigned)np;
    P.alpha = rap[0]; P.z = rap[1]; P.zrc = rap[2]; P.aux = aux;
    const unsigned gl = (unsigned)((L + 255) / 256), gr = (unsigned)((R + 255) / 256);
    prof_begin(ctx, "cairo_aux_keys");
    prof_work(ctx, 32.0 * 7 * N, 7.0 * N * 0.2, 0);
    s252::cairo_aux_keys<<<gl, 256, 0, ctx->stream>>>(P, keys.p, idx.p, okeys.p, key_or.p);
    LAUNCH_CHECK(ctx);
    // sort_columns_by_memory_address (air.rs:529-533): stable sort by address; offsets_sorted.sort() (air.rs:684-689)
    unsigned long long kor = 0;
    CU(ctx, cudaMemcpyAsync(&kor, key_or.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    unsigned key_bits = 0;
    while (key_bits < 64 && (kor >> key_bits)) ++key_bits;
    int which = 0, owhich = 0;
    TRY(radix_sort<unsigned long long>(ctx, keys.p, keys2.p, idx.p, idx2.p, L, key_bits, hist.p, &which));
    TRY(radix_sort<unsigned short>(ctx, okeys.p, okeys2.p, nullptr, nullptr, R, 16, hist.p, &owhich));
    const unsigned* sorted_idx = which ? idx2.p : idx.p;
    const unsigned short* sorted_off = owhich ? okeys2.p : okeys.p;
    prof_begin(ctx, "cairo_aux_terms");
    prof_work(ctx, 32.0 * 8 * L, 2.0 * L, 0);
    s252::cairo_aux_terms<<<gl, 256, 0, ctx->stream>>>(P, sorted_idx, num.p, den.p);
    LAUNCH_CHECK(ctx);
    prof_begin(ctx, "cairo_aux_rc_terms");
    prof_work(ctx, 32.0 * 4 * R, 1.0 * R, 0);
    s252::cairo_aux_rc_terms<<<gr, 256, 0, ctx->stream>>>(P, sorted_off, rnum.p, rden.p);
    LAUNCH_CHECK(ctx);
    TRY(scan_mul(ctx, num.p, L, false));
    TRY(scan_mul(ctx, den.p, L, true));
    TRY(scan_mul(ctx, rnum.p, R, false));
    TRY(scan_mul(ctx, rden.p, R, true));
    fe tot[2];
    CU(ctx, cudaMemcpyAsync(&tot[0], den.p, sizeof(fe), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaMemcpyAsync(&tot[1], rden.p, sizeof(fe), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    if (H::is_zero(tot[0]) || H::is_zero(tot[1])) FAIL(ctx, S252_ERR_INVALID, "a permutation-argument challenge collides with a trace value");
    prof_begin(ctx, "cairo_aux_finish");
    prof_work(ctx, 32.0 * 3 * L, 2.0 * L, 0);
    s252::cairo_aux_finish<<<gl, 256, 0, ctx->stream>>>(aux, N, 11, 4, num.p, den.p, H::inv(tot[0]));
    LAUNCH_CHECK(ctx);
    prof_begin(ctx, "cairo_aux_finish");
    prof_work(ctx, 32.0 * 3 * R, 2.0 * R, 0);
    s252::cairo_aux_finish<<<gr, 256, 0, ctx->stream>>>(aux, N, 15, 3, rnum.p, rden.p, H::inv(tot[1]));
    LAUNCH_CHECK(ctx);
    CU(ctx, cudaStreamSynchronize(ctx->stream));   // the host staging vectors go out of scope
    return S252_OK;
}

// interpolate_and_commit (prover.rs:126-159) from a column-major LW table in (pinned) host memory.  The
// upload is pipelined with the transforms: group g+1 of columns crosses PCIe on the copy stream while
// group g is converted, interpolated and extended, so only the first group's upload is exposed.
static int commit_from_host_columns(s252_ctx* ctx, const s252_fe* cols_lw, size_t N, unsigned c, size_t blowup, uint64_t coset_offset,
                                    bool keep_trace, s252_commit** out, uint8_t root[32], bool with_tree = true) {
    if (!is_pow2(N) || c == 0) FAIL(ctx, S252_ERR_INVALID, "FFTError: trace length %zu is not a power of two", N);
    if (!is_pow2(blowup) || blowup > MAX_COSETS) FAIL(ctx, S252_ERR_INVALID, "blowup factor %zu must be a power of two <= %u", blowup, MAX_COSETS);
    if (coset_offset == 0) FAIL(ctx, S252_ERR_INVALID, "coset offset must be non-zero");
    const size_t M = N * blowup;
    const std::vector<unsigned> lo = upload_groups(c, true);        // short first and last groups: the only exposed upload / transforms
    const unsigned K = (unsigned)lo.size() - 1;
    s252_commit* cm = new s252_commit();
    cm->ctx = ctx; cm->n_cols = c; cm->n_rows = M; cm->n_coeffs = N;
    std::vector<cudaEvent_t> ev(K, nullptr);
    cudaEvent_t start = nullptr;
    int rc = [&]() -> int {
        Tmp<fe> staged(ctx), cols(ctx);
        TRY(dalloc(ctx, &staged.p, N * c));
        TRY(dalloc(ctx, &cols.p, N * c));
        TRY(dalloc(ctx, &cm->coeffs, N * c));
        TRY(dalloc(ctx, &cm->lde, M * c));
        // the staging block may be a recycled one: the copy stream must not overtake work already queued on the compute stream
        CU(ctx, cudaEventCreateWithFlags(&start, cudaEventDisableTiming));
        CU(ctx, cudaEventRecord(start, ctx->stream));
        CU(ctx, cudaStreamWaitEvent(ctx->copy_stream, start, 0));
        for (unsigned g = 0; g < K; ++g) {
            const size_t off = (size_t)lo[g] * N, cnt = (size_t)(lo[g + 1] - lo[g]) * N;
            CU(ctx, cudaMemcpyAsync(staged.p + off, cols_lw + off, cnt * sizeof(fe), cudaMemcpyHostToDevice, ctx->copy_stream));
            CU(ctx, cudaEventCreateWithFlags(&ev[g], cudaEventDisableTiming));
            CU(ctx, cudaEventRecord(ev[g], ctx->copy_stream));
        }
        Xform I;
        I.logn = ilog2(N);
        I.inverse = true;
        for (unsigned g = 0; g < K; ++g) {
            const unsigned cg = lo[g + 1] - lo[g];
            const size_t off = (size_t)lo[g] * N;
            CU(ctx, cudaStreamWaitEvent(ctx->stream, ev[g], 0));
            TRY(convert_lw_to_internal(ctx, staged.p + off, cols.p + off, (size_t)cg * N));
            TRY(run_ntt(ctx, I, cols.p + off, N, false, cm->coeffs + off, N, false, cg));                 // compute_trace_polys
            TRY(evaluate_cosets(ctx, cm->coeffs + off, N, false, ilog2(N), (unsigned)blowup, H::from_u64(coset_offset),
                                cm->lde + (size_t)lo[g] * M, M, false, cg));                              // compute_lde_trace_evaluations
        }
        if (with_tree) {
            TRY(dalloc(ctx, &cm->nodes, 4 * (2 * M - 1)));
            TRY(build_tree(ctx, cm->lde, M, c, M, cm->nodes));                                           // batch_commit
            TRY(fetch_root(ctx, cm->nodes, root));
        } else {
            CU(ctx, cudaStreamSynchronize(ctx->stream));
        }
        if (keep_trace) { cm->trace = cols.p; cols.p = nullptr; }
        return S252_OK;
    }();
    if (rc != S252_OK) cudaStreamSynchronize(ctx->copy_stream);
    for (auto e : ev) if (e) cudaEventDestroy(e);
    if (start) cudaEventDestroy(start);
    if (rc != S252_OK) { commit_free(cm); return rc; }
    *out = cm;
    return S252_OK;
}

extern "C" int s252_cairo_round1(s252_ctx* ctx, const s252_cairo_trace* trace, size_t blowup, uint64_t coset_offset,
                                 s252_transcript* transcript, s252_commit** main_out, s252_commit** aux_out, s252_fe rap_out[3]) {
    NVTX_RANGE("s252_cairo_round1");
    if (!ctx || !trace || !transcript || !main_out || !aux_out || !rap_out) return S252_ERR_INVALID;
    *main_out = *aux_out = nullptr;
    CU(ctx, cudaSetDevice(ctx->device));
    const size_t N = trace->n_rows;
    uint8_t root[32];
    s252_commit* mainc = nullptr;
    s252_cairo_trace_pin(trace);   // first use only; a failure just leaves the table pageable
    StageTimer R1;
    TRY(commit_from_host_columns(ctx, trace->cols.data(), N, (unsigned)trace->n_cols, blowup, coset_offset, true, &mainc, root));
    R1.mark("main_commit");
    transcript->append(root, 32);                                   // prover.rs:151
    fe rap[3];
    for (int k = 0; k < 3; ++k) { rap[k] = transcript->to_field(); H::to_lw(rap[k], rap_out[k].limbs); }   // air.rs:731-737
    s252_commit* auxc = new s252_commit();
    auxc->ctx = ctx; auxc->n_cols = s252::CAIRO_AUX_COLS; auxc->n_rows = N * blowup; auxc->n_coeffs = N;
    int rc = [&]() -> int {
        TRY(dalloc(ctx, &auxc->trace, N * s252::CAIRO_AUX_COLS));
        TRY(cairo_build_aux(ctx, mainc->trace, 0, N, trace->pi, rap, auxc->trace));
        R1.mark("aux_build");
        TRY(lde_from_cols(ctx, auxc->trace, N, s252::CAIRO_AUX_COLS, blowup, coset_offset, true, auxc, root));
        R1.mark("aux_commit");
        return S252_OK;
    }();
    g_cairo_round1 = R1.json + "}";
    if (rc != S252_OK) { commit_free(mainc); commit_free(auxc); return rc; }
    transcript->append(root, 32);
    *main_out = mainc;
    *aux_out = auxc;
    return S252_OK;
}
extern "C" int s252_commit_read_trace(s252_commit* c, size_t col, s252_fe* out) {
    s252_ctx* ctx = c->ctx;
    CU(ctx, cudaSetDevice(ctx->device));
    if (!c->trace) FAIL(ctx, S252_ERR_INVALID, "this handle keeps no trace evaluations");
    if (col >= c->n_cols) FAIL(ctx, S252_ERR_RANGE, "column %zu out of range", col);
    return read_internal_as_lw(ctx, c->trace + col * c->n_coeffs, c->n_coeffs, out);
}

// CairoAIR::new (air.rs:594-625)
static const uint8_t CAIRO_DEGREES[50] = {2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3,
                                          2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1};

// ConstraintEvaluator::evaluate for CairoAIR (evaluator.rs:40-262) into evals[M] (device, internal format);
// ba/bb: boundary alphas/betas (8), ta/tb: transition alphas/betas (49, or 50 with the range-check builtin).
// a block of consecutive LDE rows of the two round-1 tables (the whole coset on one GPU)
struct CairoRowBlock {
    const fe *main, *aux, *hmain, *haux;      // column-major blocks and their `blowup`-row halos (see CairoEval)
    size_t stride, hstride, row0, rows;
};
static int cairo_eval_constraints_rows(s252_ctx* ctx, const s252_cairo_trace* trace, const CairoRowBlock& B, const fe rap[3],
                                       const fe* ba, const fe* bb, const fe* ta, const fe* tb, size_t blowup, uint64_t coset_offset,
                                       fe* evals);
static int cairo_eval_constraints(s252_ctx* ctx, const s252_cairo_trace* trace, const s252_commit* mainc, const s252_commit* auxc,
                                  const fe rap[3], const fe* ba, const fe* bb, const fe* ta, const fe* tb, size_t blowup,
                                  uint64_t coset_offset, fe* evals) {
    const size_t M = trace->n_rows * blowup;
    if (mainc->n_rows != M || auxc->n_rows != M || auxc->n_cols != s252::CAIRO_AUX_COLS || mainc->n_cols != trace->n_cols)
        FAIL(ctx, S252_ERR_INVALID, "round-1 handles do not match the trace");
    const CairoRowBlock B{mainc->lde, auxc->lde, mainc->lde, auxc->lde, M, M, 0, M};
    return cairo_eval_constraints_rows(ctx, trace, B, rap, ba, bb, ta, tb, blowup, coset_offset, evals);
}
static int cairo_eval_constraints_rows(s252_ctx* ctx, const s252_cairo_trace* trace, const CairoRowBlock& B, const fe rap[3],
                                       const fe* ba, const fe* bb, const fe* ta, const fe* tb, size_t blowup, uint64_t coset_offset,
                                       fe* evals) {
    const CA::PublicInputs& pi = trace->pi;
    const size_t N = trace->n_rows, M = N * blowup, b = blowup;
    if (B.rows == 0 || B.row0 + B.rows > M || B.stride < B.rows || B.hstride < b) FAIL(ctx, S252_ERR_INVALID, "bad row block");
    if (!pi.has_rc) FAIL(ctx, S252_ERR_INVALID, "public inputs carry no range-check bounds (build_main_trace sets them)");
    if (pi.num_steps == 0 || pi.num_steps > N) FAIL(ctx, S252_ERR_INVALID, "num_steps out of range");
    const bool has_rc = trace->n_cols > CA::MAIN_COLS;
    const unsigned mc = (unsigned)trace->n_cols;
    const int nt = has_rc ? 50 : 49;
    fe g, w;
    H::primitive_root(ilog2(N), &g);
    H::primitive_root(ilog2(M), &w);
    // CairoAIR::boundary_constraints (air.rs:777-849)
    struct BC { unsigned col; uint64_t step; fe value; };
    fe perm_final;
    {
        fe prod = H::one();
        for (auto& kv : pi.public_memory) prod = H::mul(prod, H::sub(rap[1], H::add(H::from_u64(kv.first), H::mul(rap[0], kv.second))));
        if (H::is_zero(prod)) FAIL(ctx, S252_ERR_INVALID, "z_memory collides with the public memory");
        perm_final = H::mul(H::pow_u64(rap[1], pi.public_memory.size()), H::inv(prod));
    }
    const BC bcs[8] = {{CA::FRAME_PC, 0, H::from_u64(pi.pc_init)}, {CA::FRAME_AP, 0, H::from_u64(pi.ap_init)},
                       {CA::FRAME_PC, pi.num_steps - 1, H::from_u64(pi.pc_final)}, {CA::FRAME_AP, pi.num_steps - 1, H::from_u64(pi.ap_final)},
                       {mc + 14, N - 1, perm_final}, {mc + 17, N - 1, H::one()},
                       {mc + 0, 0, H::from_u64(pi.range_check_min)}, {mc + 2, N - 1, H::from_u64(pi.range_check_max)}};
    // per-residue tables: x^N takes `blowup` values over the coset (x = h w^i, i mod blowup = r)
    std::vector<fe> tabs((8 + (size_t)nt) * b);
    {
        const fe hn = H::pow_u64(H::from_u64(coset_offset), N), wn = H::pow_u64(w, N);
        const fe ginv = H::inv(g);
        fe xn = hn;
        for (size_t r = 0; r < b; ++r) {
            const fe d = H::sub(xn, H::one());
            if (H::is_zero(d)) FAIL(ctx, S252_ERR_INVALID, "the LDE coset meets the trace domain");
            const fe zinv = H::inv(d), x2n = H::sqr(xn);
            for (int k = 0; k < 8; ++k)
                tabs[k * b + r] = H::mul(H::pow_u64(ginv, bcs[k].step), H::add(H::mul(ba[k], xn), bb[k]));
            for (int k = 0; k < nt; ++k) {
                const fe adj = CAIRO_DEGREES[k] == 1 ? x2n : CAIRO_DEGREES[k] == 2 ? xn : H::one();
                tabs[(8 + k) * b + r] = H::mul(H::add(H::mul(ta[k], adj), tb[k]), zinv);
            }
            xn = H::mul(xn, wn);
        }
    }
    s252::CairoEval E{};
    E.main = B.main; E.aux = B.aux; E.m = M; E.blowup = (unsigned)b; E.main_cols = mc; E.has_rc = has_rc;
    E.hmain = B.hmain; E.haux = B.haux; E.stride = B.stride; E.hstride = B.hstride; E.row0 = B.row0; E.rows = B.rows;
    E.alpha = rap[0]; E.z = rap[1]; E.zrc = rap[2];
    E.g_last = H::pow_u64(g, N - 1);
    E.two = H::from_u64(2); E.b15 = H::from_u64(1ull << 15); E.b16 = H::from_u64(1ull << 16); E.b32 = H::from_u64(1ull << 32); E.b48 = H::from_u64(1ull << 48);
    E.nb = 8;
    for (int k = 0; k < 8; ++k) { E.bcol[k] = bcs[k].col; E.bshift[k] = (b * bcs[k].step) % M; E.bval[k] = bcs[k].value; }
    TRY(get_coset_tables(ctx, M, coset_offset, &E.dom, &E.T));
    Tmp<fe> dtabs(ctx);
    TRY(dalloc(ctx, &dtabs.p, tabs.size()));
    CU(ctx, cudaMemcpyAsync(dtabs.p, tabs.data(), tabs.size() * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream));
    E.bcoef = dtabs.p; E.tcoef = dtabs.p + 8 * b;
    E.out = evals;
    const unsigned grid = (unsigned)((B.rows + s252::CAIRO_EVAL_THREADS - 1) / s252::CAIRO_EVAL_THREADS);
    prof_begin(ctx, "cairo_constraints_kernel<0>");
    prof_work(ctx, 32.0 * B.rows * 29, 53.0 * B.rows, 0);
    s252::cairo_constraints_kernel<0><<<grid, s252::CAIRO_EVAL_THREADS, 0, ctx->stream>>>(E);
    LAUNCH_CHECK(ctx);
    prof_begin(ctx, "cairo_constraints_kernel<1>");
    prof_work(ctx, 32.0 * B.rows * 29, 42.0 * B.rows, 0);
    s252::cairo_constraints_kernel<1><<<grid, s252::CAIRO_EVAL_THREADS, 0, ctx->stream>>>(E);
    LAUNCH_CHECK(ctx);
    prof_begin(ctx, "cairo_constraints_kernel<2>");
    prof_work(ctx, 32.0 * B.rows * 44, 70.0 * B.rows, 0);
    s252::cairo_constraints_kernel<2><<<grid, s252::CAIRO_EVAL_THREADS, 0, ctx->stream>>>(E);
    LAUNCH_CHECK(ctx);
    CU(ctx, cudaStreamSynchronize(ctx->stream));   // the coefficient tables (host vector, device temporary) go out of scope
    return S252_OK;
}

extern "C" int s252_cairo_constraint_evaluations(s252_ctx* ctx, const s252_cairo_trace* trace, s252_commit* mainc, s252_commit* auxc,
                                                 const s252_fe rap_lw[3], const s252_fe* boundary_coeffs, const s252_fe* transition_coeffs,
                                                 size_t blowup, uint64_t coset_offset, s252_fe* out) {
    NVTX_RANGE("s252_cairo_constraint_evaluations");
    if (!ctx || !trace || !mainc || !auxc || !rap_lw || !boundary_coeffs || !transition_coeffs || !out) return S252_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    const int nt = trace->n_cols > CA::MAIN_COLS ? 50 : 49;
    fe rap[3], ba[8], bb[8], ta[50], tb[50];
    for (int k = 0; k < 3; ++k) rap[k] = H::from_lw(rap_lw[k].limbs);
    for (int k = 0; k < 8; ++k) { ba[k] = H::from_lw(boundary_coeffs[2 * k].limbs); bb[k] = H::from_lw(boundary_coeffs[2 * k + 1].limbs); }
    for (int k = 0; k < nt; ++k) { ta[k] = H::from_lw(transition_coeffs[2 * k].limbs); tb[k] = H::from_lw(transition_coeffs[2 * k + 1].limbs); }
    const size_t M = trace->n_rows * blowup;
    Tmp<fe> evals(ctx);
    TRY(dalloc(ctx, &evals.p, M));
    TRY(cairo_eval_constraints(ctx, trace, mainc, auxc, rap, ba, bb, ta, tb, blowup, coset_offset, evals.p));
    return read_internal_as_lw(ctx, evals.p, M, out);
}

// round_2_compute_composition_polynomial after the constraint evaluation (prover.rs:246-283): H from its M
// evaluations on the coset (interpolate_offset_fft, evaluation_table.rs:27-33), even/odd split, LDE of H1 and
// H2, batch_commit.  The handle keeps the H1, H2 coefficients.
static int cairo_composition_commit(s252_ctx* ctx, const fe* evals, size_t N, size_t blowup, uint64_t coset_offset,
                                    s252_commit** out, uint8_t root[32], bool with_tree = true) {
    const size_t M = N * blowup;
    s252_commit* cm = new s252_commit();
    cm->ctx = ctx; cm->n_cols = 2; cm->n_rows = M; cm->n_coeffs = N;
    int rc = [&]() -> int {
        Tmp<fe> hco(ctx);
        Tmp<unsigned> flag(ctx);
        TRY(dalloc(ctx, &hco.p, M));
        Xform X;
        X.logn = ilog2(M);
        X.inverse = true;
        X.has_offset_scale = true;
        X.oscale_base = H::inv(H::from_u64(coset_offset));
        TRY(run_ntt(ctx, X, evals, M, false, hco.p, M, false, 1));
        // even_odd_decomposition (prover.rs:252)
        TRY(dalloc(ctx, &cm->coeffs, 2 * N));
        TRY(dalloc(ctx, &flag.p, 1));
        CU(ctx, cudaMemsetAsync(flag.p, 0, 4, ctx->stream));
        prof_begin(ctx, "split_even_odd");
        prof_work(ctx, 32.0 * (M + 2 * N), 0, 0);
        s252::split_even_odd<<<(unsigned)((M + 255) / 256), 256, 0, ctx->stream>>>(hco.p, M, N, cm->coeffs, flag.p);
        LAUNCH_CHECK(ctx);
        unsigned over = 0;
        CU(ctx, cudaMemcpyAsync(&over, flag.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        if (over) FAIL(ctx, S252_ERR_INVALID, "the composition polynomial exceeds its degree bound: the trace does not satisfy the Cairo AIR");
        // evaluate_polynomial_on_lde_domain(H1), (H2) + batch_commit (prover.rs:254-276)
        TRY(dalloc(ctx, &cm->lde, 2 * M));
        TRY(evaluate_cosets(ctx, cm->coeffs, N, false, ilog2(N), (unsigned)blowup, H::from_u64(coset_offset), cm->lde, M, false, 2));
        if (with_tree) {
            TRY(dalloc(ctx, &cm->nodes, 4 * (2 * M - 1)));
            TRY(build_tree(ctx, cm->lde, M, 2, M, cm->nodes));
            TRY(fetch_root(ctx, cm->nodes, root));
        }
        return S252_OK;
    }();
    if (rc != S252_OK) { commit_free(cm); return rc; }
    *out = cm;
    return S252_OK;
}

extern "C" int s252_cairo_round2(s252_ctx* ctx, const s252_cairo_trace* trace, s252_commit* mainc, s252_commit* auxc,
                                 const s252_fe rap_lw[3], size_t blowup, uint64_t coset_offset, s252_transcript* transcript,
                                 s252_commit** composition_out) {
    NVTX_RANGE("s252_cairo_round2");
    if (!ctx || !trace || !mainc || !auxc || !rap_lw || !transcript || !composition_out) return S252_ERR_INVALID;
    *composition_out = nullptr;
    CU(ctx, cudaSetDevice(ctx->device));
    const size_t N = trace->n_rows, M = N * blowup;
    const int nt = trace->n_cols > CA::MAIN_COLS ? 50 : 49;
    fe rap[3];
    for (int k = 0; k < 3; ++k) rap[k] = H::from_lw(rap_lw[k].limbs);
    // <<<< challenges (prover.rs:598-626): boundary alphas, boundary betas, transition alphas, transition betas
    fe ba[8], bb[8], ta[50], tb[50];
    for (int k = 0; k < 8; ++k) ba[k] = transcript->to_field();
    for (int k = 0; k < 8; ++k) bb[k] = transcript->to_field();
    for (int k = 0; k < nt; ++k) ta[k] = transcript->to_field();
    for (int k = 0; k < nt; ++k) tb[k] = transcript->to_field();
    Tmp<fe> evals(ctx);
    TRY(dalloc(ctx, &evals.p, M));
    TRY(cairo_eval_constraints(ctx, trace, mainc, auxc, rap, ba, bb, ta, tb, blowup, coset_offset, evals.p));
    uint8_t root[32];
    TRY(cairo_composition_commit(ctx, evals.p, N, blowup, coset_offset, composition_out, root));
    transcript->append(root, 32);                                   // prover.rs:635
    return S252_OK;
}

// --------------------------------------------------------------------------------------------
// Building blocks of a proof sharded over several GPUs (lambdaworks_cairo_prover_b200/cairo_distributed.py):
// the same kernels as s252_cairo_prove, on this rank's columns (LDE) or on this rank's block of LDE rows.
extern "C" const s252_fe* s252_cairo_trace_columns(const s252_cairo_trace* t) { return t->cols.data(); }

extern "C" int s252_lde_host_columns(s252_ctx* ctx, const s252_fe* cols_lw, size_t n_rows, size_t n_cols, size_t blowup,
                                     uint64_t coset_offset, int keep_trace, s252_commit** out) {
    NVTX_RANGE("s252_lde_host_columns");
    if (!ctx || !cols_lw || !out) return S252_ERR_INVALID;
    *out = nullptr;
    CU(ctx, cudaSetDevice(ctx->device));
    return commit_from_host_columns(ctx, cols_lw, n_rows, (unsigned)n_cols, blowup, coset_offset, keep_trace != 0, out, nullptr, false);
}
extern "C" const void* s252_commit_device_trace(const s252_commit* c) { return c->trace; }
extern "C" int s252_lde_device_columns(s252_ctx* ctx, const void* cols, size_t n_rows, size_t n_cols, size_t blowup,
                                       uint64_t coset_offset, s252_commit** out) {
    NVTX_RANGE("s252_lde_device_columns");
    if (!ctx || !cols || !out) return S252_ERR_INVALID;
    *out = nullptr;
    CU(ctx, cudaSetDevice(ctx->device));
    if (!is_pow2(n_rows) || n_cols == 0 || !is_pow2(blowup) || blowup > MAX_COSETS || coset_offset == 0) FAIL(ctx, S252_ERR_INVALID, "bad LDE shape");
    s252_commit* cm = new s252_commit();
    cm->ctx = ctx; cm->n_cols = n_cols; cm->n_rows = n_rows * blowup; cm->n_coeffs = n_rows;
    const int rc = lde_from_cols(ctx, reinterpret_cast<const fe*>(cols), n_rows, (unsigned)n_cols, blowup, coset_offset, false, cm, nullptr);
    if (rc != S252_OK) { commit_free(cm); return rc; }
    *out = cm;
    return S252_OK;
}
// build_auxiliary_trace on this device from the 11 main-trace columns it reads (pc .. off_op1, uploaded from
// the handle's pinned column-major table): *aux_out = device buffer [18][n_rows] (internal format; free it
// with s252_device_free).
extern "C" int s252_cairo_aux_trace_device(s252_ctx* ctx, const s252_cairo_trace* trace, const s252_fe rap_lw[3], const void* prefetched,
                                           int prefetched_internal, void** aux_out) {
    NVTX_RANGE("s252_cairo_aux_trace_device");
    if (!ctx || !trace || !rap_lw || !aux_out) return S252_ERR_INVALID;
    *aux_out = nullptr;
    CU(ctx, cudaSetDevice(ctx->device));
    const size_t N = trace->n_rows;
    fe rap[3];
    for (int k = 0; k < 3; ++k) rap[k] = H::from_lw(rap_lw[k].limbs);
    s252_cairo_trace_pin(trace);
    Tmp<fe> staged(ctx), cols(ctx);
    fe* aux = nullptr;
    const fe* src = reinterpret_cast<const fe*>(prefetched);
    const fe* in_cols = src;
    if (!src || !prefetched_internal) {
        TRY(dalloc(ctx, &cols.p, 11 * N));
        if (!src) {
            TRY(dalloc(ctx, &staged.p, 11 * N));
            CU(ctx, cudaMemcpyAsync(staged.p, trace->cols.data() + (size_t)s252::CAIRO_PC * N, 11 * N * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream));
            src = staged.p;
        }
        TRY(convert_lw_to_internal(ctx, src, cols.p, 11 * N));
        in_cols = cols.p;
    }
    TRY(dalloc(ctx, &aux, (size_t)s252::CAIRO_AUX_COLS * N));
    const int rc = cairo_build_aux(ctx, in_cols, s252::CAIRO_PC, N, trace->pi, rap, aux);
    if (rc != S252_OK) { dfree(ctx, aux); return rc; }
    *aux_out = aux;
    return S252_OK;
}
extern "C" int s252_cairo_constraints_rows(s252_ctx* ctx, const s252_cairo_trace* trace, const void* main_block, const void* aux_block,
                                           size_t stride, size_t row0, size_t rows, const void* main_halo, const void* aux_halo,
                                           size_t halo_stride, const s252_fe rap_lw[3], const s252_fe* boundary_coeffs,
                                           const s252_fe* transition_coeffs, size_t blowup, uint64_t coset_offset, void* evals_out) {
    NVTX_RANGE("s252_cairo_constraints_rows");
    if (!ctx || !trace || !main_block || !aux_block || !main_halo || !aux_halo || !rap_lw || !boundary_coeffs || !transition_coeffs || !evals_out)
        return S252_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    const int nt = trace->n_cols > CA::MAIN_COLS ? 50 : 49;
    fe rap[3], ba[8], bb[8], ta[50], tb[50];
    for (int k = 0; k < 3; ++k) rap[k] = H::from_lw(rap_lw[k].limbs);
    for (int k = 0; k < 8; ++k) { ba[k] = H::from_lw(boundary_coeffs[2 * k].limbs); bb[k] = H::from_lw(boundary_coeffs[2 * k + 1].limbs); }
    for (int k = 0; k < nt; ++k) { ta[k] = H::from_lw(transition_coeffs[2 * k].limbs); tb[k] = H::from_lw(transition_coeffs[2 * k + 1].limbs); }
    const CairoRowBlock B{reinterpret_cast<const fe*>(main_block), reinterpret_cast<const fe*>(aux_block), reinterpret_cast<const fe*>(main_halo),
                          reinterpret_cast<const fe*>(aux_halo), stride, halo_stride, row0, rows};
    return cairo_eval_constraints_rows(ctx, trace, B, rap, ba, bb, ta, tb, blowup, coset_offset, reinterpret_cast<fe*>(evals_out));
}
extern "C" int s252_cairo_composition_commit(s252_ctx* ctx, const void* evals, size_t n_rows, size_t blowup, uint64_t coset_offset,
                                             s252_commit** out, uint8_t root[32]) {
    NVTX_RANGE("s252_cairo_composition_commit");
    if (!ctx || !evals || !out || !root) return S252_ERR_INVALID;
    *out = nullptr;
    CU(ctx, cudaSetDevice(ctx->device));
    if (!is_pow2(n_rows) || !is_pow2(blowup) || coset_offset == 0) FAIL(ctx, S252_ERR_INVALID, "bad composition shape");
    return cairo_composition_commit(ctx, reinterpret_cast<const fe*>(evals), n_rows, blowup, coset_offset, out, root);
}
// The same without the tree: H1/H2 coefficients + their LDE (a rank of a sharded proof hashes only its own rows).
extern "C" int s252_cairo_composition_lde(s252_ctx* ctx, const void* evals, size_t n_rows, size_t blowup, uint64_t coset_offset,
                                          s252_commit** out) {
    NVTX_RANGE("s252_cairo_composition_lde");
    if (!ctx || !evals || !out) return S252_ERR_INVALID;
    *out = nullptr;
    CU(ctx, cudaSetDevice(ctx->device));
    if (!is_pow2(n_rows) || !is_pow2(blowup) || coset_offset == 0) FAIL(ctx, S252_ERR_INVALID, "bad composition shape");
    return cairo_composition_commit(ctx, reinterpret_cast<const fe*>(evals), n_rows, blowup, coset_offset, out, nullptr, false);
}
extern "C" int s252_deep_rows(s252_ctx* ctx, const void* const* tables, const size_t* strides, const size_t* n_cols, size_t n_tables,
                              size_t row0, size_t rows, size_t lde_rows, size_t trace_rows, const s252_fe* z,
                              const uint64_t* transition_offsets, size_t n_offsets, const s252_fe* trace_ood, const s252_fe* h1_z2,
                              const s252_fe* h2_z2, const s252_fe* gamma, const s252_fe* gamma_p, const s252_fe* trace_gammas,
                              uint64_t coset_offset, void* out) {
    NVTX_RANGE("s252_deep_rows");
    if (!ctx || !tables || !strides || !n_cols || !z || !transition_offsets || !trace_ood || !h1_z2 || !h2_z2 || !gamma || !gamma_p ||
        !trace_gammas || !out || n_tables < 2 || n_tables > (size_t)s252::DEEP_MAX_TABLES)
        return S252_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    DeepTables T{};
    for (size_t i = 0; i < n_tables; ++i) { T.cols[i] = reinterpret_cast<const fe*>(tables[i]); T.strides[i] = strides[i]; T.ncols[i] = (unsigned)n_cols[i]; }
    T.ntables = (unsigned)n_tables;
    return deep_evaluate_rows(ctx, T, row0, rows, lde_rows, trace_rows, z, transition_offsets, n_offsets, trace_ood, h1_z2, h2_z2, gamma,
                              gamma_p, trace_gammas, coset_offset, reinterpret_cast<fe*>(out));
}

// StarkProof::serialize helpers (proof/stark.rs:161-218; frame.rs:86-105; fri_decommit.rs:19-45; utils.rs:6-13)
namespace {
struct ByteSink {
    std::vector<uint8_t> b;
    void u64(uint64_t v) { uint8_t t[8]; put_u64_be(t, v); b.insert(b.end(), t, t + 8); }
    void raw(const uint8_t* p, size_t n) { b.insert(b.end(), p, p + n); }
    void felt(const s252_fe& v) { uint8_t t[32]; H::to_bytes_be(H::from_lw(v.limbs), t); raw(t, 32); }
    void path(const uint8_t* p, size_t depth) { u64(depth); raw(p, depth * 32); }
    void blob(const ByteSink& o) { u64(o.b.size()); raw(o.b.data(), o.b.size()); }
};
}  // namespace

// StarkProof::serialize (proof/stark.rs:161-218) from the pieces of a proof: ood = the frame (2 x cols, row-major), FRI values and
// paths as s252_fri_query returns them ([Q][layers] values, [Q][layers][depth] digests), openings as s252_commit_open returns them.
static void serialize_stark_proof(ByteSink& S, size_t N, const uint8_t* root_main, const uint8_t* root_aux, const uint8_t* root_comp,
                                  const s252_fe* ood, size_t cols, const s252_fe hz[2], size_t layers, const uint8_t* fri_roots,
                                  const s252_fe& last, size_t Q, size_t depth, const s252_fe* evs, const s252_fe* ev, const uint8_t* pas,
                                  const uint8_t* pa, const s252_fe* comp_rows, const uint8_t* comp_paths, const s252_fe* main_rows,
                                  size_t main_cols, const uint8_t* main_paths, const s252_fe* aux_rows, size_t aux_cols,
                                  const uint8_t* aux_paths, uint64_t nonce) {
    S.b.reserve(4096 + Q * (2 * layers * (depth + 2) + 3 * (depth + 1) + cols + 8) * 32);
    S.u64(N);
    S.u64(2);
    S.raw(root_main, 32);
    S.raw(root_aux, 32);
    {
        ByteSink F;
        F.u64(2 * cols); F.u64(32);
        for (size_t i = 0; i < 2 * cols; ++i) F.felt(ood[i]);
        F.u64(cols);
        S.blob(F);
    }
    S.raw(root_comp, 32);
    S.u64(32); S.felt(hz[0]); S.felt(hz[1]);
    S.u64(layers); S.raw(fri_roots, 32 * layers);
    S.felt(last);
    S.u64(Q);
    for (size_t q = 0; q < Q; ++q) {
        ByteSink D;
        D.u64(layers);
        for (size_t k = 0; k < layers; ++k) D.path(pas + (q * layers + k) * depth * 32, depth - k);
        D.u64(32);
        D.u64(layers);
        for (size_t k = 0; k < layers; ++k) D.felt(evs[q * layers + k]);
        D.u64(layers);
        for (size_t k = 0; k < layers; ++k) D.felt(ev[q * layers + k]);
        D.u64(layers);
        for (size_t k = 0; k < layers; ++k) D.path(pa + (q * layers + k) * depth * 32, depth - k);
        S.blob(D);
    }
    S.u64(Q);
    for (size_t q = 0; q < Q; ++q) {
        ByteSink O;
        O.path(comp_paths + q * depth * 32, depth);
        O.u64(32); O.felt(comp_rows[2 * q]); O.felt(comp_rows[2 * q + 1]);
        O.u64(2);
        O.path(main_paths + q * depth * 32, depth);
        O.path(aux_paths + q * depth * 32, depth);
        O.u64(cols);
        for (size_t j = 0; j < main_cols; ++j) O.felt(main_rows[q * main_cols + j]);
        for (size_t j = 0; j < aux_cols; ++j) O.felt(aux_rows[q * aux_cols + j]);
        S.blob(O);
    }
    S.u64(nonce);
}
// The same for a caller that assembled the pieces itself (the sharded prover gathers them from several GPUs).
extern "C" int s252_cairo_serialize_proof(size_t trace_rows, const uint8_t* root_main, const uint8_t* root_aux, const uint8_t* root_comp,
                                          const s252_fe* ood, size_t cols, const s252_fe* hz, size_t layers, const uint8_t* fri_roots,
                                          const s252_fe* last, size_t n_queries, size_t depth, const s252_fe* evs, const s252_fe* ev,
                                          const uint8_t* pas, const uint8_t* pa, const s252_fe* comp_rows, const uint8_t* comp_paths,
                                          const s252_fe* main_rows, size_t main_cols, const uint8_t* main_paths, const s252_fe* aux_rows,
                                          size_t aux_cols, const uint8_t* aux_paths, uint64_t nonce, uint8_t** proof_out, size_t* proof_len) {
    if (!root_main || !root_aux || !root_comp || !ood || !hz || !last || !proof_out || !proof_len || main_cols + aux_cols != cols ||
        (layers && !fri_roots) || (n_queries && (!evs || !ev || !pas || !pa || !comp_rows || !comp_paths || !main_rows || !main_paths || !aux_rows || !aux_paths)))
        CAIRO_FAIL(S252_ERR_INVALID, "s252_cairo_serialize_proof: bad arguments");
    try {
        ByteSink S;
        serialize_stark_proof(S, trace_rows, root_main, root_aux, root_comp, ood, cols, hz, layers, fri_roots, *last, n_queries, depth, evs, ev,
                              pas, pa, comp_rows, comp_paths, main_rows, main_cols, main_paths, aux_rows, aux_cols, aux_paths, nonce);
        uint8_t* outp = (uint8_t*)std::malloc(S.b.size() ? S.b.size() : 1);
        if (!outp) CAIRO_FAIL(S252_ERR_INVALID, "out of host memory");
        std::memcpy(outp, S.b.data(), S.b.size());
        *proof_out = outp;
        *proof_len = S.b.size();
    } catch (const std::exception& e) {
        CAIRO_FAIL(S252_ERR_INVALID, std::string("s252_cairo_serialize_proof: ") + e.what());
    }
    return S252_OK;
}

extern "C" int s252_cairo_prove(s252_ctx* ctx, const s252_cairo_trace* trace, size_t blowup, size_t n_queries, uint64_t coset_offset,
                                uint8_t grinding_factor, uint8_t** proof_out, size_t* proof_len) {
    NVTX_RANGE("s252_cairo_prove");
    if (!ctx || !trace || !proof_out || !proof_len) return S252_ERR_INVALID;
    *proof_out = nullptr; *proof_len = 0;
    CU(ctx, cudaSetDevice(ctx->device));
    const size_t N = trace->n_rows, M = N * blowup;
    if (!is_pow2(N) || N < 2) FAIL(ctx, S252_ERR_INVALID, "trace length %zu is not a power of two", N);
    s252_transcript t;                                              // round_0_transcript_initialization
    s252_commit *mainc = nullptr, *auxc = nullptr, *comp = nullptr;
    s252_fri* fri = nullptr;
    s252_fe rap[3];
    StageTimer ST;
    auto body = [&]() -> int {
        TRY(s252_cairo_round1(ctx, trace, blowup, coset_offset, &t, &mainc, &auxc, rap));
        ST.mark("round1");
        TRY(s252_cairo_round2(ctx, trace, mainc, auxc, rap, blowup, coset_offset, &t, &comp));
        ST.mark("round2");
        // ---- round 3 (prover.rs:650-690): z, H1(z^2), H2(z^2), t_j(z g^k)
        nvtxRangePushA("round3_ood");
        fe g;
        H::primitive_root(ilog2(N), &g);
        const fe hinv = H::inv(H::from_u64(coset_offset));
        fe z;
        for (;;) {                                                  // sample_z_ood (transcript.rs:53-70)
            z = t.to_field();
            if (!H::eq(H::pow_u64(H::mul(z, hinv), M), H::one()) && !H::eq(H::pow_u64(z, N), H::one())) break;
        }
        const size_t cols = mainc->n_cols + auxc->n_cols;
        s252_fe pts[2], z2, zlw, hz[2];
        H::to_lw(z, pts[0].limbs); H::to_lw(H::mul(z, g), pts[1].limbs); H::to_lw(H::sqr(z), z2.limbs);
        zlw = pts[0];
        std::vector<s252_fe> ood(2 * cols);
        TRY(s252_commit_evaluate_at(mainc, pts, 2, ood.data(), cols, 0));
        TRY(s252_commit_evaluate_at(auxc, pts, 2, ood.data(), cols, mainc->n_cols));
        TRY(s252_commit_evaluate_at(comp, &z2, 1, hz, 2, 0));
        uint8_t be[32];
        for (int k = 0; k < 2; ++k) { H::to_bytes_be(H::from_lw(hz[k].limbs), be); t.append(be, 32); }
        for (auto& v : ood) { H::to_bytes_be(H::from_lw(v.limbs), be); t.append(be, 32); }
        nvtxRangePop();
        ST.mark("round3");
        // ---- round 4 (prover.rs:327-404)
        s252_fe gamma, gamma_p;
        H::to_lw(t.to_field(), gamma.limbs); H::to_lw(t.to_field(), gamma_p.limbs);
        std::vector<s252_fe> tg(2 * cols);
        for (auto& v : tg) H::to_lw(t.to_field(), v.limbs);
        const size_t layers = ilog2(N);
        const uint64_t offs[2] = {0, 1};
        s252_commit* tcs[2] = {mainc, auxc};
        s252_fe last;
        std::vector<uint8_t> fri_roots(32 * std::max<size_t>(layers, 1));
        TRY(s252_fri_commit_phase_deep(ctx, layers, tcs, 2, comp, &zlw, offs, 2, ood.data(), &hz[0], &hz[1], &gamma, &gamma_p, tg.data(), &t,
                                       coset_offset, &fri, &last, fri_roots.data()));
        ST.mark("deep_fri");
        uint8_t challenge[32];
        t.challenge(challenge);
        uint64_t nonce = 0;
        TRY(s252_generate_nonce_with_grinding(ctx, challenge, grinding_factor, 0, &nonce));
        uint8_t nb[8];
        put_u64_be(nb, nonce);
        t.append(nb, 8);
        ST.mark("grinding");
        const size_t Q = layers ? n_queries : 0;                   // fri_query_phase returns nothing without layers (fri/mod.rs:83)
        std::vector<uint64_t> iotas(Q);
        for (auto& i : iotas) i = t.to_usize() % M;
        const size_t depth = ilog2(M);
        std::vector<s252_fe> ev(Q * layers), evs(Q * layers), main_rows(Q * mainc->n_cols), aux_rows(Q * auxc->n_cols), comp_rows(Q * 2);
        std::vector<uint8_t> pa(Q * layers * depth * 32), pas(Q * layers * depth * 32), main_paths(Q * depth * 32), aux_paths(Q * depth * 32),
            comp_paths(Q * depth * 32);
        if (Q) {
            TRY(s252_fri_query(fri, iotas.data(), Q, ev.data(), evs.data(), pa.data(), pas.data(), depth));
            TRY(s252_commit_open(mainc, iotas.data(), Q, main_rows.data(), main_paths.data()));
            TRY(s252_commit_open(auxc, iotas.data(), Q, aux_rows.data(), aux_paths.data()));
            TRY(s252_commit_open(comp, iotas.data(), Q, comp_rows.data(), comp_paths.data()));
        }
        ST.mark("openings");
        // ---- StarkProof::serialize
        uint8_t r_main[32], r_aux[32], r_comp[32];
        TRY(s252_commit_root(mainc, r_main));
        TRY(s252_commit_root(auxc, r_aux));
        TRY(s252_commit_root(comp, r_comp));
        ByteSink S;
        serialize_stark_proof(S, N, r_main, r_aux, r_comp, ood.data(), cols, hz, layers, fri_roots.data(), last, Q, depth, evs.data(), ev.data(),
                              pas.data(), pa.data(), comp_rows.data(), comp_paths.data(), main_rows.data(), mainc->n_cols, main_paths.data(),
                              aux_rows.data(), auxc->n_cols, aux_paths.data(), nonce);
        uint8_t* outp = (uint8_t*)std::malloc(S.b.size());
        if (!outp) FAIL(ctx, S252_ERR_INVALID, "out of host memory");
        std::memcpy(outp, S.b.data(), S.b.size());
        *proof_out = outp;
        *proof_len = S.b.size();
        ST.mark("serialize");
        return S252_OK;
    };
    int rc;
    try {
        rc = body();
    } catch (const std::exception& e) {           // host allocation failure: nothing may unwind across the C boundary
        ctx->err = std::string("s252_cairo_prove: ") + e.what();
        rc = S252_ERR_INVALID;
    }
    if (fri) s252_fri_destroy(fri);
    commit_free(mainc); commit_free(auxc); commit_free(comp);
    ST.mark("free");
    g_cairo_stages = ST.json + ", \"round1_detail\": " + g_cairo_round1 + "}";
    return rc;
}
extern "C" void s252_cairo_proof_free(uint8_t* proof) { std::free(proof); }
