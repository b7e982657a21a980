// runtime.cu -- host runtime + C ABI (include/stark252_b200.h) over the sm_100a kernels.
//
// One context = one device + one stream + a cache of twiddle tables + a device-memory arena that
// recycles the multi-GB LDE / scratch buffers of consecutive commits without going back to the
// driver.
#include <algorithm>
#include <array>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "../../include/stark252_b200.h"
#include "cairo.cuh"
#include "commit.cuh"
#include "deep.cuh"
#include "fe.cuh"
#include "host_field.hpp"
#include "keccak.cuh"
#include "microbench.cuh"
#include "ntt.cuh"

using s252::fe;
namespace H = s252::host;

// --------------------------------------------------------------------------------------------
struct s252_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // host->device prefetch of the next trace, overlapping compute
    cudaEvent_t copy_event = nullptr;
    unsigned* ticket = nullptr;           // device counter of merkle_finish (zero between launches)
    std::string err;
    uint64_t launches = 0;
    std::map<std::string, fe*> tables;   // cached twiddle tables (device)
    size_t table_bytes = 0;
    unsigned max_logl = s252::NTT_MAX_LOGL;
    // Device memory arena: freed blocks are kept and handed back to later requests of (nearly) the
    // same size.  Everything runs on one stream, so reuse needs no synchronisation, and a prover
    // that commits traces of the same shape over and over never goes back to the driver.
    std::multimap<size_t, void*> arena_free;
    std::map<void*, size_t> arena_size;
    size_t arena_bytes = 0;
    // per-kernel device timing (s252_ctx_profile*): events around every launch
    bool prof = false;
    // work = algorithmic HBM bytes (one read + one write of the data the kernel must touch),
    // field multiplications and Keccak-f permutations of the launch ("work accounting")
    struct Work { double bytes = 0, muls = 0, perms = 0; };
    struct Pending { const char* name; cudaEvent_t e0, e1; Work w; };
    struct Acc { uint64_t launches = 0; double ms = 0; Work w; };
    std::vector<Pending> prof_pending;
    std::vector<cudaEvent_t> prof_free;
    std::map<std::string, Acc> prof_acc;
    cudaEvent_t prof_e0 = nullptr;
    const char* prof_name = nullptr;
    Work prof_work;
};

static cudaEvent_t prof_event(s252_ctx* ctx) {
    if (!ctx->prof_free.empty()) { cudaEvent_t e = ctx->prof_free.back(); ctx->prof_free.pop_back(); return e; }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}
static inline void prof_begin(s252_ctx* ctx, const char* name) {
    if (!ctx->prof) return;
    ctx->prof_e0 = prof_event(ctx);
    ctx->prof_name = name;
    ctx->prof_work = s252_ctx::Work();
    cudaEventRecord(ctx->prof_e0, ctx->stream);
}
static inline void prof_work(s252_ctx* ctx, double bytes, double muls, double perms) {
    ctx->prof_work.bytes = bytes;
    ctx->prof_work.muls = muls;
    ctx->prof_work.perms = perms;
}
static inline void prof_end(s252_ctx* ctx) {
    if (!ctx->prof || !ctx->prof_e0) return;
    cudaEvent_t e1 = prof_event(ctx);
    cudaEventRecord(e1, ctx->stream);
    ctx->prof_pending.push_back({ctx->prof_name, ctx->prof_e0, e1, ctx->prof_work});
    ctx->prof_e0 = nullptr;
}
static void prof_drain(s252_ctx* ctx) {
    cudaStreamSynchronize(ctx->stream);
    for (auto& pnd : ctx->prof_pending) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, pnd.e0, pnd.e1) == cudaSuccess) {
            auto& acc = ctx->prof_acc[pnd.name];
            acc.launches += 1;
            acc.ms += ms;
            acc.w.bytes += pnd.w.bytes;
            acc.w.muls += pnd.w.muls;
            acc.w.perms += pnd.w.perms;
        }
        ctx->prof_free.push_back(pnd.e0);
        ctx->prof_free.push_back(pnd.e1);
    }
    ctx->prof_pending.clear();
}

struct s252_commit {
    s252_ctx* ctx = nullptr;
    size_t n_cols = 0, n_rows = 0, n_coeffs = 0;
    fe* coeffs = nullptr;       // [n_cols][n_coeffs]
    fe* lde = nullptr;          // [n_cols][n_rows]
    uint64_t* nodes = nullptr;  // [(2*n_rows-1)][4]
    fe* trace = nullptr;        // [n_cols][n_coeffs] trace evaluations, kept only for the Cairo prover (aux-trace build)
    bool owns_lde = true;       // false: the columns belong to the caller (s252_commit_device_columns_inplace)
};

struct FriLayerDev {
    size_t size = 0;
    fe* evals = nullptr;
    uint64_t* nodes = nullptr;
};
struct s252_fri {
    s252_ctx* ctx = nullptr;
    size_t domain_size = 0;
    std::vector<FriLayerDev> layers;
    fe h0 = s252::fe_one();     // coset offset of layer 0 (layer-by-layer interface)
};

// NVTX range per entry point / prover stage (visible in Nsight Systems; a no-op without a profiler attached)
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};
#define NVTX_RANGE(name) NvtxRange _nvtx_range_(name)

#define FAIL(ctx, code, ...)                                     \
    do {                                                         \
        char _b[512];                                            \
        std::snprintf(_b, sizeof _b, __VA_ARGS__);               \
        (ctx)->err = _b;                                         \
        return (code);                                           \
    } while (0)
#define CU(ctx, call)                                                                                  \
    do {                                                                                               \
        cudaError_t _e = (call);                                                                       \
        if (_e != cudaSuccess) FAIL(ctx, S252_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(_e), \
                                    __FILE__, __LINE__);                                               \
    } while (0)
#define TRY(expr)                  \
    do {                           \
        int _rc = (expr);          \
        if (_rc != S252_OK) return _rc; \
    } while (0)
#define LAUNCH_CHECK(ctx)           \
    do {                            \
        (ctx)->launches++;          \
        prof_end(ctx);              \
        CU(ctx, cudaGetLastError()); \
    } while (0)

static inline bool is_pow2(size_t n) { return n && !(n & (n - 1)); }
static inline unsigned ilog2(size_t n) { unsigned k = 0; while (((size_t)1 << k) < n) ++k; return k; }
static inline size_t next_pow2(size_t n) { size_t r = 1; while (r < n) r <<= 1; return r; }

static void arena_trim(s252_ctx* ctx) {
    cudaStreamSynchronize(ctx->stream);
    for (auto& kv : ctx->arena_free) {
        cudaFree(kv.second);
        ctx->arena_bytes -= kv.first;
        ctx->arena_size.erase(kv.second);
    }
    ctx->arena_free.clear();
}
static int arena_alloc(s252_ctx* ctx, void** p, size_t bytes) {
    bytes = (bytes + 511) & ~(size_t)511;
    if (bytes == 0) bytes = 512;
    auto it = ctx->arena_free.lower_bound(bytes);
    if (it != ctx->arena_free.end() && it->first <= bytes + bytes / 8) {
        *p = it->second;
        ctx->arena_free.erase(it);
        return S252_OK;
    }
    cudaError_t e = cudaMalloc(p, bytes);
    if (e == cudaErrorMemoryAllocation) {
        cudaGetLastError();
        arena_trim(ctx);
        e = cudaMalloc(p, bytes);
    }
    if (e != cudaSuccess) {
        char b[256];
        std::snprintf(b, sizeof b, "cudaMalloc(%zu bytes): %s", bytes, cudaGetErrorString(e));
        ctx->err = b;
        cudaGetLastError();
        return S252_ERR_CUDA;
    }
    ctx->arena_size[*p] = bytes;
    ctx->arena_bytes += bytes;
    return S252_OK;
}
template <typename T>
static int dalloc(s252_ctx* ctx, T** p, size_t count) {
    *p = nullptr;
    return arena_alloc(ctx, (void**)p, count * sizeof(T));
}
template <typename T>
static void dfree(s252_ctx* ctx, T* p) {
    if (!p) return;
    auto it = ctx->arena_size.find((void*)p);
    if (it == ctx->arena_size.end()) return;
    ctx->arena_free.emplace(it->second, (void*)p);
}
// RAII for temporaries
template <typename T>
struct Tmp {
    s252_ctx* ctx;
    T* p = nullptr;
    explicit Tmp(s252_ctx* c) : ctx(c) {}
    ~Tmp() { dfree(ctx, p); }
    Tmp(const Tmp&) = delete;
    Tmp& operator=(const Tmp&) = delete;
};

// --------------------------------------------------------------------------------------------
// context
extern "C" int s252_ctx_create(int device, s252_ctx** out) {
    if (!out) return S252_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0 || device < 0 || device >= count) return S252_ERR_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return S252_ERR_CUDA;
    s252_ctx* ctx = new s252_ctx();
    ctx->device = device;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return S252_ERR_CUDA; }
    if (cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->copy_event, cudaEventDisableTiming) != cudaSuccess) { delete ctx; return S252_ERR_CUDA; }
    if (cudaMalloc((void**)&ctx->ticket, 64) != cudaSuccess || cudaMemset(ctx->ticket, 0, 64) != cudaSuccess) { delete ctx; return S252_ERR_CUDA; }
    cudaFuncSetAttribute(s252::ntt_pass_strided, cudaFuncAttributeMaxDynamicSharedMemorySize, s252::NTT_TILE * 32);
    cudaFuncSetAttribute(s252::ntt_pass_final, cudaFuncAttributeMaxDynamicSharedMemorySize, s252::NTT_TILE * 32);
    if (const char* e = std::getenv("S252_MAX_LOGL")) {
        int v = std::atoi(e);
        if (v >= 4 && v <= s252::NTT_MAX_LOGL) ctx->max_logl = (unsigned)v;
    }
    *out = ctx;
    return S252_OK;
}
extern "C" void s252_ctx_destroy(s252_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    arena_trim(ctx);
    for (auto& kv : ctx->arena_size) cudaFree(kv.first);   // blocks still owned by live handles
    for (auto& kv : ctx->tables) cudaFree(kv.second);
    cudaFree(ctx->ticket);
    cudaStreamSynchronize(ctx->copy_stream);
    cudaEventDestroy(ctx->copy_event);
    cudaStreamDestroy(ctx->copy_stream);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}
extern "C" const char* s252_last_error(const s252_ctx* ctx) { return ctx ? ctx->err.c_str() : "no context"; }
extern "C" int s252_ctx_synchronize(s252_ctx* ctx) {
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return S252_OK;
}
extern "C" void* s252_ctx_stream(s252_ctx* ctx) { return (void*)ctx->stream; }
extern "C" uint64_t s252_ctx_launch_count(const s252_ctx* ctx) { return ctx->launches; }
extern "C" int s252_ctx_trim(s252_ctx* ctx) {
    if (!ctx) return S252_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    arena_trim(ctx);
    return S252_OK;
}
extern "C" int s252_ctx_profile(s252_ctx* ctx, int enable) {
    if (!ctx) return S252_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    prof_drain(ctx);
    ctx->prof = enable != 0;
    if (enable == 2) ctx->prof_acc.clear();
    return S252_OK;
}
extern "C" int s252_ctx_profile_read(s252_ctx* ctx, char* buf, size_t cap) {
    if (!ctx || !buf || cap == 0) return S252_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    prof_drain(ctx);
    std::string out = "{";
    bool first = true;
    for (auto& kv : ctx->prof_acc) {
        char line[384];
        std::snprintf(line, sizeof line, "%s\"%s\": {\"launches\": %llu, \"ms\": %.6f, \"bytes\": %.0f, \"muls\": %.0f, \"perms\": %.0f}",
                      first ? "" : ", ", kv.first.c_str(), (unsigned long long)kv.second.launches, kv.second.ms,
                      kv.second.w.bytes, kv.second.w.muls, kv.second.w.perms);
        out += line;
        first = false;
    }
    out += "}";
    if (out.size() + 1 > cap) FAIL(ctx, S252_ERR_INVALID, "profile buffer too small (%zu needed)", out.size() + 1);
    std::memcpy(buf, out.c_str(), out.size() + 1);
    return S252_OK;
}
extern "C" int s252_device_alloc(s252_ctx* ctx, size_t bytes, void** out) {
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaMalloc(out, bytes ? bytes : 1));
    return S252_OK;
}
extern "C" int s252_device_free(s252_ctx* ctx, void* ptr) {
    if (!ctx) return S252_ERR_INVALID;
    if (!ptr) return S252_OK;
    // buffers the library handed out from its arena (s252_cairo_aux_trace_device, ..) go back to the arena: reuse is ordered on the
    // context's stream, so neither a synchronisation nor a cudaFree (a device-wide barrier, milliseconds for a GB) is needed
    if (ctx->arena_size.find(ptr) != ctx->arena_size.end()) {
        dfree(ctx, ptr);
        return S252_OK;
    }
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    CU(ctx, cudaFree(ptr));
    return S252_OK;
}
extern "C" int s252_copy_to_device(s252_ctx* ctx, void* dst, const void* src, size_t bytes) {
    CU(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return S252_OK;
}
extern "C" int s252_host_register(void* ptr, size_t bytes) {
    if (!ptr || !bytes) return S252_ERR_INVALID;
    if (cudaHostRegister(ptr, bytes, cudaHostRegisterDefault) != cudaSuccess) { cudaGetLastError(); return S252_ERR_CUDA; }
    return S252_OK;
}
extern "C" int s252_host_unregister(void* ptr) {
    if (!ptr) return S252_ERR_INVALID;
    if (cudaHostUnregister(ptr) != cudaSuccess) { cudaGetLastError(); return S252_ERR_CUDA; }
    return S252_OK;
}
// Strided host -> device copy on the compute stream (a slab of a matrix: `height` runs of `width` bytes).
extern "C" int s252_copy_2d_to_device(s252_ctx* ctx, void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t width,
                                      size_t height) {
    if (!ctx || !dst || !src) return S252_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, width, height, cudaMemcpyHostToDevice, ctx->stream));
    return S252_OK;
}
// Prefetch: enqueue a host->device copy on the context's copy stream (returns immediately; use
// pinned host memory for a true asynchronous DMA).  The destination must not be in use by work
// already queued on the compute stream.
extern "C" int s252_copy_to_device_async(s252_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (!ctx || !dst || !src) return S252_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
    return S252_OK;
}
// Make everything issued on the compute stream from now on wait for the prefetches issued so far.
extern "C" int s252_copy_stream_wait(s252_ctx* ctx) {
    if (!ctx) return S252_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaEventRecord(ctx->copy_event, ctx->copy_stream));
    CU(ctx, cudaStreamWaitEvent(ctx->stream, ctx->copy_event, 0));
    return S252_OK;
}
extern "C" int s252_copy_to_host(s252_ctx* ctx, void* dst, const void* src, size_t bytes) {
    CU(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return S252_OK;
}

// --------------------------------------------------------------------------------------------
// twiddle tables
static std::string fe_key(const fe& a) {
    char b[80];
    std::snprintf(b, sizeof b, "%08x%08x%08x%08x%08x%08x%08x%08x", a.l[7], a.l[6], a.l[5], a.l[4], a.l[3], a.l[2], a.l[1], a.l[0]);
    return b;
}
static int table_alloc(s252_ctx* ctx, const std::string& key, size_t count, fe** out, bool* fresh) {
    auto it = ctx->tables.find(key);
    if (it != ctx->tables.end()) { *out = it->second; *fresh = false; return S252_OK; }
    fe* p = nullptr;
    CU(ctx, cudaMalloc((void**)&p, count * sizeof(fe)));
    ctx->tables[key] = p;
    ctx->table_bytes += count * sizeof(fe);
    *out = p;
    *fresh = true;
    return S252_OK;
}
static int get_level_table(s252_ctx* ctx, unsigned logL, unsigned ncosets, const fe& wL, const fe& shift,
                           const fe& step, const fe** out) {
    std::string key = "lvl:" + std::to_string(logL) + ":" + std::to_string(ncosets) + ":" + fe_key(wL) + ":" +
                      fe_key(shift) + ":" + fe_key(step);
    fe* t; bool fresh;
    const size_t count = ((size_t)1 << logL) * ncosets;
    TRY(table_alloc(ctx, key, count, &t, &fresh));
    if (fresh) {
        prof_begin(ctx, "gen_level_twiddles");
        s252::gen_level_twiddles<<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(t, logL, ncosets, wL, shift, step);
        LAUNCH_CHECK(ctx);
    }
    *out = t;
    return S252_OK;
}
static int get_pass_table(s252_ctx* ctx, unsigned logL, unsigned logInner, unsigned ncosets, const fe& wS,
                          const fe& shift, const fe& step, const fe& scale, const fe** out) {
    std::string key = "ptw:" + std::to_string(logL) + ":" + std::to_string(logInner) + ":" + std::to_string(ncosets) +
                      ":" + fe_key(wS) + ":" + fe_key(shift) + ":" + fe_key(step) + ":" + fe_key(scale);
    fe* t; bool fresh;
    const size_t count = ((size_t)1 << (logL + logInner)) * ncosets;
    TRY(table_alloc(ctx, key, count, &t, &fresh));
    if (fresh) {
        prof_begin(ctx, "gen_pass_twiddles");
        s252::gen_pass_twiddles<<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(t, logL, logInner, ncosets, wS, shift, step, scale);
        LAUNCH_CHECK(ctx);
    }
    *out = t;
    return S252_OK;
}
static int get_power_table(s252_ctx* ctx, size_t n, const fe& base, const fe& scale, const fe** out) {
    std::string key = "pow:" + std::to_string(n) + ":" + fe_key(base) + ":" + fe_key(scale);
    fe* t; bool fresh;
    TRY(table_alloc(ctx, key, n, &t, &fresh));
    if (fresh) {
        prof_begin(ctx, "gen_powers");
        s252::gen_powers<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(t, n, base, scale);
        LAUNCH_CHECK(ctx);
    }
    *out = t;
    return S252_OK;
}

// --------------------------------------------------------------------------------------------
// One batched transform over `ncols` columns of N = 2^logn elements (internal format unless in_lw).
//   forward:  out[(k*ncosets + c)] = sum_n in[n] * (shift*step^c)^n * w_N^(n k)      (natural order)
//   inverse:  out[k] = oscale[k] * (1/N) * sum_n in[n] * w_N^(-n k)
struct Xform {
    unsigned logn = 0;
    bool inverse = false;
    unsigned ncosets = 1;
    fe shift = s252::fe_one();       // coset shift of coset 0 (forward only)
    fe step = s252::fe_one();        // ratio between consecutive coset shifts (forward only)
    bool has_offset_scale = false;   // inverse only: multiply output k by oscale_base^k
    fe oscale_base = s252::fe_one();
};

// A transform shared by `parts` GPUs: phase 0 = the first pass on this GPU's range of inner positions (into zbuf),
// phase 1 = the remaining passes on this GPU's range of first-digit rows (from zbuf, after the all-to-all).
struct NttShare {
    int phase = -1;              // -1: the whole transform on this GPU
    unsigned part = 0, parts = 1;
    fe* zbuf = nullptr;          // [ncols][ncosets][N], owned by the caller (it crosses the exchange)
    unsigned* log_l1 = nullptr;  // out: the first digit (the exchange geometry: rows of 2^l1, N / 2^l1 inner positions)
};
static int run_ntt(s252_ctx* ctx, const Xform& X, const fe* in, size_t in_col_stride, bool in_lw, fe* out,
                   size_t out_col_stride, bool out_lw, unsigned ncols, const NttShare* share = nullptr) {
    const unsigned logn = X.logn;
    const size_t N = (size_t)1 << logn;
    const unsigned maxl = ctx->max_logl;
    if (ncols == 0) return S252_OK;
    if (X.ncosets == 0 || !is_pow2(X.ncosets)) FAIL(ctx, S252_ERR_INVALID, "coset count must be a power of two");
    fe wN;
    if (!H::primitive_root(logn, &wN)) FAIL(ctx, S252_ERR_INVALID, "no root of unity of order 2^%u", logn);
    auto root = [&](unsigned lg) { fe w; H::primitive_root(lg, &w); return X.inverse ? H::inv(w) : w; };
    const fe n_inv = X.inverse ? H::inv(H::from_u64((uint64_t)N)) : H::one();

    s252::NttPass P{};
    P.ncosets = X.ncosets;
    P.ncols = ncols;
    P.ostride = X.ncosets;
    const unsigned smem = s252::NTT_TILE * 32;

    // digits
    unsigned npass = logn <= maxl ? 1 : (logn <= 2 * maxl ? 2 : 3);
    if (logn > 3 * maxl) FAIL(ctx, S252_ERR_INVALID, "transform of size 2^%u is not supported", logn);

    const fe* oscale = nullptr;
    if (X.inverse && (X.has_offset_scale || npass == 1)) {
        TRY(get_power_table(ctx, N, X.has_offset_scale ? X.oscale_base : H::one(), npass == 1 ? n_inv : H::one(), &oscale));
    }

    const bool shared = share && share->phase >= 0;
    if (shared && (npass == 1 || !share->zbuf || share->parts == 0 || !is_pow2(share->parts) || share->part >= share->parts))
        FAIL(ctx, S252_ERR_INVALID, "a shared transform needs at least two passes (size > 2^%u) and a power-of-two number of parts", maxl);
    if (npass == 1) {
        const fe* lvl;
        TRY(get_level_table(ctx, logn, X.ncosets, root(logn), X.shift, X.step, &lvl));
        unsigned logT = s252::NTT_TILE_LOG - logn;
        if (logT > 5) logT = 5;
        while (logT > 0 && (1u << logT) > ncols) --logT;
        P.in = in; P.out = out; P.lvl = lvl; P.ptw = nullptr; P.oscale = oscale;
        P.in_col_stride = in_col_stride; P.out_col_stride = out_col_stride;
        P.in_coset_stride = 0; P.out_coset_stride = 0;
        P.logL = logn; P.logT = logT; P.logN1 = 0; P.logN2 = 0;
        P.lvl_per_coset = X.ncosets > 1; P.rows_are_cols = 1; P.in_lw = in_lw; P.out_lw = out_lw;
        P.first_unit = (X.ncosets == 1 && H::eq(X.shift, H::one())) ? 1 : 0;
        const unsigned tiles = (ncols + (1u << logT) - 1) >> logT;
        prof_begin(ctx, "ntt_pass_final");
        prof_work(ctx, 32.0 * N * ncols * (1 + X.ncosets), 0.5 * N * logn * X.ncosets * ncols + (oscale ? (double)N * ncols : 0.0), 0);
        s252::ntt_pass_final<<<dim3(tiles * X.ncosets, 1), s252::NTT_THREADS, smem, ctx->stream>>>(P);
        LAUNCH_CHECK(ctx);
        return S252_OK;
    }
    if (in_lw) FAIL(ctx, S252_ERR_INVALID, "internal: LW input only on the single-pass path");

    unsigned l1, l2 = 0, l3;
    if (npass == 2) {
        l3 = (logn + 1) / 2;
        l1 = logn - l3;
    } else {
        l3 = (logn + 2) / 3;
        l2 = (logn - l3 + 1) / 2;
        l1 = logn - l3 - l2;
    }
    const unsigned lparts = shared ? ilog2(share->parts) : 0;
    if (shared) {
        if (share->log_l1) *share->log_l1 = l1;
        if (l1 < lparts || logn - l1 < lparts) FAIL(ctx, S252_ERR_INVALID, "too many parts for a transform of size 2^%u", logn);
        if (share->phase == 2) return S252_OK;       // geometry query only
    }
    // scratch: [col][coset][N]
    Tmp<fe> Z(ctx);
    if (shared) Z.p = share->zbuf; else TRY(dalloc(ctx, &Z.p, (size_t)ncols * X.ncosets * N));
    struct Release { Tmp<fe>& z; bool keep; ~Release() { if (keep) z.p = nullptr; } } release{Z, shared};   // the caller's buffer is not ours to free
    P.block0 = 0; P.part_g0 = 0; P.part_gn = 0;

    // pass A1: L = 2^l1 over stride 2^(logn-l1)
    if (!shared || share->phase == 0) {
        const unsigned logInner = logn - l1;
        const fe shiftL = H::pow_u64(X.shift, (uint64_t)1 << logInner);
        const fe stepL = H::pow_u64(X.step, (uint64_t)1 << logInner);
        const fe *lvl, *ptw;
        TRY(get_level_table(ctx, l1, X.ncosets, root(l1), shiftL, stepL, &lvl));
        TRY(get_pass_table(ctx, l1, logInner, X.ncosets, root(logn), X.shift, X.step, n_inv, &ptw));
        unsigned logT = std::min(s252::NTT_TILE_LOG - l1, logInner - lparts);
        if (logT > 5) logT = 5;
        P.in = in; P.out = Z.p; P.lvl = lvl; P.ptw = ptw; P.oscale = nullptr;
        P.in_col_stride = in_col_stride; P.out_col_stride = (size_t)X.ncosets * N;
        P.in_coset_stride = 0; P.out_coset_stride = N;
        P.logL = l1; P.logT = logT; P.logInner = logInner; P.logOuter = 0;
        P.lvl_per_coset = X.ncosets > 1; P.ptw_per_coset = X.ncosets > 1; P.rows_are_cols = 0; P.in_lw = 0; P.out_lw = 0;
        P.first_unit = (X.ncosets == 1 && H::eq(X.shift, H::one())) ? 1 : 0;
        size_t tiles = (size_t)1 << (logInner - logT);
        if (shared) {            // this GPU's inner positions [part * inner / parts, ..): a contiguous range of tiles
            tiles >>= lparts;
            P.block0 = (unsigned)(tiles * share->part * X.ncosets * ncols);
        }
        prof_begin(ctx, "ntt_pass_strided");
        prof_work(ctx, 32.0 * N * ncols * (1 + X.ncosets) / (shared ? share->parts : 1), (0.5 * l1 + 1.0) * N * X.ncosets * ncols / (shared ? share->parts : 1), 0);
        s252::ntt_pass_strided<<<(unsigned)(tiles * X.ncosets * ncols), s252::NTT_THREADS, smem, ctx->stream>>>(P);
        LAUNCH_CHECK(ctx);
        P.block0 = 0;
    }
    if (shared && share->phase == 0) return S252_OK;
    if (npass == 3) {
        // pass A2: in place on Z, view [2^l1][2^l2][2^l3]
        const fe *lvl, *ptw;
        TRY(get_level_table(ctx, l2, 1, root(l2), H::one(), H::one(), &lvl));
        TRY(get_pass_table(ctx, l2, l3, 1, root(l2 + l3), H::one(), H::one(), H::one(), &ptw));
        unsigned logT = std::min(s252::NTT_TILE_LOG - l2, l3);
        if (logT > 5) logT = 5;
        P.in = Z.p; P.out = Z.p; P.lvl = lvl; P.ptw = ptw; P.oscale = nullptr;
        P.in_col_stride = (size_t)X.ncosets * N; P.out_col_stride = (size_t)X.ncosets * N;
        P.in_coset_stride = N; P.out_coset_stride = N;
        P.logL = l2; P.logT = logT; P.logInner = l3; P.logOuter = l1;
        P.lvl_per_coset = 0; P.ptw_per_coset = 0; P.first_unit = 1;
        size_t tiles = (size_t)1 << (l1 + l3 - logT);
        if (shared) {            // this GPU's first-digit rows [part * 2^l1 / parts, ..) = a contiguous range of outer indices
            tiles >>= lparts;
            P.block0 = (unsigned)(tiles * share->part * X.ncosets * ncols);
        }
        prof_begin(ctx, "ntt_pass_strided");
        prof_work(ctx, 64.0 * N * ncols * X.ncosets / (shared ? share->parts : 1), (0.5 * l2 + 1.0) * N * X.ncosets * ncols / (shared ? share->parts : 1), 0);
        s252::ntt_pass_strided<<<(unsigned)(tiles * X.ncosets * ncols), s252::NTT_THREADS, smem, ctx->stream>>>(P);
        LAUNCH_CHECK(ctx);
        P.block0 = 0;
    }
    {
        // final pass: rows of 2^l3
        const fe* lvl;
        TRY(get_level_table(ctx, l3, 1, root(l3), H::one(), H::one(), &lvl));
        unsigned logT = std::min(s252::NTT_TILE_LOG - l3, l1 - lparts);
        if (logT > 5) logT = 5;
        P.in = Z.p; P.out = out; P.lvl = lvl; P.ptw = nullptr; P.oscale = oscale;
        P.in_col_stride = (size_t)X.ncosets * N; P.out_col_stride = out_col_stride;
        P.in_coset_stride = N; P.out_coset_stride = 0;
        P.logL = l3; P.logT = logT; P.logN1 = l1; P.logN2 = l2;
        P.lvl_per_coset = 0; P.ptw_per_coset = 0; P.rows_are_cols = 0; P.in_lw = 0; P.out_lw = out_lw; P.first_unit = 1;
        size_t tiles = (size_t)1 << (l1 + l2 - logT);
        if (shared) {            // this GPU's groups of T consecutive first-digit rows, for every k2
            P.part_gn = 1u << (l1 - logT - lparts);
            P.part_g0 = P.part_gn * share->part;
            tiles >>= lparts;
        }
        prof_begin(ctx, "ntt_pass_final");
        prof_work(ctx, 64.0 * N * ncols * X.ncosets / (shared ? share->parts : 1),
                  (0.5 * l3 * N * X.ncosets * ncols + (oscale ? (double)N * ncols : 0.0)) / (shared ? share->parts : 1), 0);
        s252::ntt_pass_final<<<(unsigned)(tiles * X.ncosets * ncols), s252::NTT_THREADS, smem, ctx->stream>>>(P);
        LAUNCH_CHECK(ctx);
        P.part_g0 = 0; P.part_gn = 0;
    }
    return S252_OK;
}

// Node levels of a tree whose 2^depth leaf digests are in place.  Wide levels go three at a time
// through merkle_subtrees (every thread busy); from level 15..17 down one launch of merkle_finish
// reaches the root (and, for a FRI layer, advances the device-side transcript chain).
static int build_tree_nodes(s252_ctx* ctx, size_t n_rows, uint64_t* nodes, s252::FriChain* chain = nullptr) {
    unsigned level = ilog2(n_rows);
    while (level > 17) {
        const unsigned lv = std::min(3u, level - 15);
        const size_t groups = (size_t)1 << (level - lv);
        const unsigned blocks = (unsigned)((groups + 127) / 128);
        const double parents = (double)((size_t)1 << level) - (double)groups;
        prof_begin(ctx, "merkle_subtrees");
        prof_work(ctx, 32.0 * ((size_t)1 << level) + 32.0 * parents, 0, parents);
        if (lv == 3) s252::merkle_subtrees<3><<<blocks, 128, 0, ctx->stream>>>(nodes, level);
        else if (lv == 2) s252::merkle_subtrees<2><<<blocks, 128, 0, ctx->stream>>>(nodes, level);
        else s252::merkle_subtrees<1><<<blocks, 128, 0, ctx->stream>>>(nodes, level);
        LAUNCH_CHECK(ctx);
        level -= lv;
    }
    if (level > 0 || chain) {
        const size_t nchildren = (size_t)1 << level;
        const unsigned blocks = level > (unsigned)s252::MERKLE_FUSED_LEVELS ? 1u << (level - s252::MERKLE_FUSED_LEVELS) : 1u;
        prof_begin(ctx, "merkle_finish");
        prof_work(ctx, 64.0 * nchildren, 0, (double)(nchildren - 1) + (chain ? 1.0 : 0.0));
        s252::merkle_finish<<<blocks, s252::MERKLE_BLOCK, 0, ctx->stream>>>(nodes, level, ctx->ticket, chain);
        LAUNCH_CHECK(ctx);
    }
    return S252_OK;
}

// Batched Merkle tree over `n_rows` rows of column-major `cols`.
static int build_tree(s252_ctx* ctx, const fe* cols, size_t col_stride, unsigned ncols, size_t n_rows, uint64_t* nodes,
                      s252::FriChain* chain = nullptr) {
    if (!is_pow2(n_rows)) FAIL(ctx, S252_ERR_INVALID, "merkle tree needs a power-of-two number of leaves (got %zu)", n_rows);
    prof_begin(ctx, "merkle_leaves");
    prof_work(ctx, 32.0 * n_rows * ncols + 32.0 * n_rows, 0.2 * n_rows * ncols, (double)n_rows * ((32 * ncols) / 136 + 1));
    s252::merkle_leaves<<<(unsigned)((n_rows + 127) / 128), 128, 0, ctx->stream>>>(cols, col_stride, ncols, n_rows,
                                                                                nodes + 4 * (n_rows - 1));
    LAUNCH_CHECK(ctx);
    return build_tree_nodes(ctx, n_rows, nodes, chain);
}

// Bring a caller buffer of `count` LW elements onto the device (no format change).
static int stage_in(s252_ctx* ctx, const s252_fe* src, size_t count, int mem, Tmp<fe>& staged, const fe** dev) {
    if (mem == S252_DEVICE) { *dev = reinterpret_cast<const fe*>(src); return S252_OK; }
    TRY(dalloc(ctx, &staged.p, count));
    CU(ctx, cudaMemcpyAsync(staged.p, src, count * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream));
    *dev = staged.p;
    return S252_OK;
}
static int stage_out(s252_ctx* ctx, const fe* dev_lw, s252_fe* dst, size_t count, int mem) {
    if (mem == S252_DEVICE) return S252_OK;   // written in place
    CU(ctx, cudaMemcpyAsync(dst, dev_lw, count * sizeof(fe), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return S252_OK;
}
static int convert_lw_to_internal(s252_ctx* ctx, const fe* in, fe* out, size_t n) {
    prof_begin(ctx, "lw_to_internal");
    s252::lw_to_internal<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(in, out, n);
    LAUNCH_CHECK(ctx);
    return S252_OK;
}

// --------------------------------------------------------------------------------------------
// FFTPoly entry points
static int interpolate_common(s252_ctx* ctx, const s252_fe* evals, size_t n, const s252_fe* offset, s252_fe* coeffs, int mem) {
    if (!ctx || !evals || !coeffs) return S252_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    if (!is_pow2(n)) FAIL(ctx, S252_ERR_INVALID, "FFTError: input length %zu is not a power of two", n);
    Tmp<fe> staged(ctx), conv(ctx), outbuf(ctx);
    const fe* din;
    TRY(stage_in(ctx, evals, n, mem, staged, &din));
    Xform X;
    X.logn = ilog2(n);
    X.inverse = true;
    if (offset) {
        const fe off = H::from_lw(offset->limbs);
        if (H::is_zero(off)) FAIL(ctx, S252_ERR_INVALID, "coset offset must be non-zero");
        X.has_offset_scale = true;
        X.oscale_base = H::inv(off);
    }
    fe* dout = reinterpret_cast<fe*>(coeffs);
    if (mem == S252_HOST) { TRY(dalloc(ctx, &outbuf.p, n)); dout = outbuf.p; }
    bool in_lw = true;
    if (X.logn > ctx->max_logl) {
        TRY(dalloc(ctx, &conv.p, n));
        TRY(convert_lw_to_internal(ctx, din, conv.p, n));
        din = conv.p;
        in_lw = false;
    }
    TRY(run_ntt(ctx, X, din, n, in_lw, dout, n, true, 1));
    TRY(stage_out(ctx, dout, coeffs, n, mem));
    return S252_OK;
}
extern "C" int s252_interpolate_fft(s252_ctx* ctx, const s252_fe* evals, size_t n, s252_fe* coeffs, int mem) {
    NVTX_RANGE("s252_interpolate_fft");
    return interpolate_common(ctx, evals, n, nullptr, coeffs, mem);
}
extern "C" int s252_interpolate_offset_fft(s252_ctx* ctx, const s252_fe* evals, size_t n, const s252_fe* offset,
                                           s252_fe* coeffs, int mem) {
    NVTX_RANGE("s252_interpolate_offset_fft");
    if (!offset) return S252_ERR_INVALID;
    return interpolate_common(ctx, evals, n, offset, coeffs, mem);
}
extern "C" size_t s252_evaluate_offset_fft_len(size_t n_coeffs, size_t blowup, size_t domain_size) {
    return next_pow2(std::max(n_coeffs, domain_size)) * blowup;
}

// Evaluate `ncols` polynomials (internal format, [ncols][n_pad] with n_pad = 2^k coefficients,
// zero padded) on the coset  offset * w_len^i, len = n_pad * ncosets.  out: [ncols][len].
static const unsigned MAX_COSETS = 64;
static int evaluate_cosets(s252_ctx* ctx, const fe* coeffs, size_t coeff_stride, bool in_lw, unsigned logn_pad,
                           unsigned ncosets, const fe& offset, fe* out, size_t out_stride, bool out_lw, unsigned ncols) {
    const size_t len = ((size_t)1 << logn_pad) * ncosets;
    fe wlen;
    if (!H::primitive_root(ilog2(len), &wlen)) FAIL(ctx, S252_ERR_INVALID, "domain of size %zu has no root of unity", len);
    Xform X;
    X.logn = logn_pad;
    X.ncosets = ncosets;
    X.shift = offset;
    X.step = wlen;
    return run_ntt(ctx, X, coeffs, coeff_stride, in_lw, out, out_stride, out_lw, ncols);
}

// Shared by evaluate_offset_fft / evaluate_polynomial_on_lde_domain / FRI layer 0: coefficients in
// LW format on the device -> evaluations on a domain of `len` points (len >= n_coeffs, power of two).
// Pads the coefficient vector so that at most MAX_COSETS cosets are needed.
static int evaluate_from_lw(s252_ctx* ctx, const fe* dcoeffs_lw, size_t n_coeffs, size_t len, const fe& offset, fe* out,
                            bool out_lw) {
    size_t n_pad = next_pow2(std::max<size_t>(n_coeffs, 1));
    if (n_pad > len) FAIL(ctx, S252_ERR_INVALID, "polynomial with %zu coefficients does not fit a domain of %zu", n_coeffs, len);
    while (len / n_pad > MAX_COSETS) n_pad <<= 1;
    Tmp<fe> pad(ctx);
    TRY(dalloc(ctx, &pad.p, n_pad));
    if (n_pad > n_coeffs) CU(ctx, cudaMemsetAsync(pad.p + n_coeffs, 0, (n_pad - n_coeffs) * sizeof(fe), ctx->stream));
    if (n_coeffs) TRY(convert_lw_to_internal(ctx, dcoeffs_lw, pad.p, n_coeffs));
    return evaluate_cosets(ctx, pad.p, n_pad, false, ilog2(n_pad), (unsigned)(len / n_pad), offset, out, len, out_lw, 1);
}

extern "C" int s252_evaluate_offset_fft(s252_ctx* ctx, const s252_fe* coeffs, size_t n_coeffs, size_t blowup,
                                        size_t domain_size, const s252_fe* offset, s252_fe* out, size_t out_capacity, int mem) {
    NVTX_RANGE("s252_evaluate_offset_fft");
    if (!ctx || !offset || !out || (!coeffs && n_coeffs)) return S252_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    if (!is_pow2(blowup)) FAIL(ctx, S252_ERR_INVALID, "blowup factor %zu is not a power of two", blowup);
    const size_t len = s252_evaluate_offset_fft_len(n_coeffs, blowup, domain_size);
    if (out_capacity < len) FAIL(ctx, S252_ERR_INVALID, "output buffer holds %zu elements, %zu needed", out_capacity, len);
    Tmp<fe> staged(ctx), outbuf(ctx);
    const fe* din = nullptr;
    if (n_coeffs) TRY(stage_in(ctx, coeffs, n_coeffs, mem, staged, &din));
    fe* dout = reinterpret_cast<fe*>(out);
    if (mem == S252_HOST) { TRY(dalloc(ctx, &outbuf.p, len)); dout = outbuf.p; }
    TRY(evaluate_from_lw(ctx, din, n_coeffs, len, H::from_lw(offset->limbs), dout, true));
    TRY(stage_out(ctx, dout, out, len, mem));
    return S252_OK;
}
extern "C" int s252_evaluate_polynomial_on_lde_domain(s252_ctx* ctx, const s252_fe* coeffs, size_t n_coeffs, size_t blowup,
                                                      size_t domain_size, const s252_fe* offset, s252_fe* out, int mem) {
    NVTX_RANGE("s252_evaluate_polynomial_on_lde_domain");
    if (!ctx || !offset || !out || (!coeffs && n_coeffs)) return S252_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    if (!is_pow2(blowup) || !is_pow2(domain_size)) FAIL(ctx, S252_ERR_INVALID, "blowup and domain size must be powers of two");
    const size_t len = s252_evaluate_offset_fft_len(n_coeffs, blowup, domain_size);
    const size_t want = domain_size * blowup;
    const size_t step = len / want;
    Tmp<fe> staged(ctx), full(ctx), outbuf(ctx);
    const fe* din = nullptr;
    if (n_coeffs) TRY(stage_in(ctx, coeffs, n_coeffs, mem, staged, &din));
    fe* dout = reinterpret_cast<fe*>(out);
    if (mem == S252_HOST) { TRY(dalloc(ctx, &outbuf.p, want)); dout = outbuf.p; }
    if (step == 1) {
        TRY(evaluate_from_lw(ctx, din, n_coeffs, len, H::from_lw(offset->limbs), dout, true));
    } else {
        // prover.rs:118-122: evaluate on the larger domain and keep every step-th point
        TRY(dalloc(ctx, &full.p, len));
        TRY(evaluate_from_lw(ctx, din, n_coeffs, len, H::from_lw(offset->limbs), full.p, true));
        prof_begin(ctx, "subsample");
        s252::subsample<<<(unsigned)((want + 255) / 256), 256, 0, ctx->stream>>>(full.p, dout, want, step);
        LAUNCH_CHECK(ctx);
    }
    TRY(stage_out(ctx, dout, out, want, mem));
    return S252_OK;
}

// One column's transform shared by `parts` GPUs (a single oversized column; four-step with one
// all-to-all).  The column is seen as a 2^l1 x (N / 2^l1) matrix (first digit x inner position):
//   phase 0   this GPU runs the first pass on ITS range of inner positions  [part * inner / parts, ..)  of `in`
//             (natural order, only that slab needs to be valid) and leaves the result in z at the same positions;
//   exchange  (caller, NCCL) every GPU collects the rows  k1 in [part * 2^l1 / parts, ..)  of z from all the others;
//   phase 1   this GPU runs the remaining passes on its rows and writes its outputs -- natural index
//             k = k1 + 2^l1 * q, hence runs of 2^l1 / parts consecutive values (times n_cosets) -- into `out`.
//   phase 2   only reports l1.
// inverse != 0: interpolate_fft (n_cosets must be 1).  Otherwise evaluate_offset_fft(n_cosets, Some(N), coset_offset):
// out[(k * n_cosets + c)].  in: N elements, z: n_cosets * N, out: n_cosets * N (device, internal format).
extern "C" int s252_ntt_shared(s252_ctx* ctx, unsigned log_n, int inverse, size_t n_cosets, uint64_t coset_offset, int phase, unsigned part,
                               unsigned parts, const void* in, void* z, void* out, unsigned* log_l1) {
    NVTX_RANGE("s252_ntt_shared");
    if (!ctx || phase < 0 || phase > 2 || (phase != 2 && (!in || !z || !out))) return S252_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    if (n_cosets == 0 || !is_pow2(n_cosets) || n_cosets > MAX_COSETS || (inverse && n_cosets != 1)) FAIL(ctx, S252_ERR_INVALID, "bad coset count");
    const size_t N = (size_t)1 << log_n;
    Xform X;
    X.logn = log_n;
    if (inverse) {
        X.inverse = true;
    } else {
        if (coset_offset == 0) FAIL(ctx, S252_ERR_INVALID, "coset offset must be non-zero");
        fe wlen;
        if (!H::primitive_root(ilog2(N * n_cosets), &wlen)) FAIL(ctx, S252_ERR_INVALID, "domain has no root of unity");
        X.ncosets = (unsigned)n_cosets;
        X.shift = H::from_u64(coset_offset);
        X.step = wlen;
    }
    NttShare sh;
    sh.phase = phase; sh.part = part; sh.parts = parts; sh.zbuf = reinterpret_cast<fe*>(z); sh.log_l1 = log_l1;
    if (phase == 2) sh.zbuf = reinterpret_cast<fe*>(ctx->ticket);       // any non-null pointer: nothing is launched
    return run_ntt(ctx, X, reinterpret_cast<const fe*>(in), N, false, reinterpret_cast<fe*>(out), N * n_cosets, false, 1, &sh);
}

extern "C" int s252_convert_elements(s252_ctx* ctx, const void* in, void* out, size_t n, int to_internal) {
    if (!ctx || !in || !out) return S252_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    if (n == 0) return S252_OK;
    prof_begin(ctx, to_internal ? "lw_to_internal" : "internal_to_lw");
    if (to_internal) s252::lw_to_internal<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(reinterpret_cast<const fe*>(in), reinterpret_cast<fe*>(out), n);
    else s252::internal_to_lw<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(reinterpret_cast<const fe*>(in), reinterpret_cast<fe*>(out), n);
    LAUNCH_CHECK(ctx);
    return S252_OK;
}

// --------------------------------------------------------------------------------------------
// commits
static void commit_free(s252_commit* c) {
    if (!c) return;
    dfree(c->ctx, c->coeffs);
    if (c->owns_lde) dfree(c->ctx, c->lde);
    dfree(c->ctx, c->nodes);
    dfree(c->ctx, c->trace);
    delete c;
}
static int fetch_root(s252_ctx* ctx, const uint64_t* nodes, uint8_t root[32]) {
    CU(ctx, cudaMemcpyAsync(root, nodes, 32, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return S252_OK;
}

// compute_trace_polys + compute_lde_trace_evaluations of interpolate_and_commit; with_tree adds
// batch_commit.  Without the tree the handle is what one rank of a column-sharded commit holds
// before the exchange (DESIGN.md "multi-GPU").
// compute_trace_polys + compute_lde_trace_evaluations (+ batch_commit) over trace columns that are already
// column-major on the device in the internal format: cols[j * N + i].
static int lde_from_cols(s252_ctx* ctx, const fe* cols, size_t N, unsigned c, size_t blowup, uint64_t coset_offset, bool with_tree,
                         s252_commit* cm, uint8_t root[32]) {
    const size_t M = N * blowup;
    // compute_trace_polys: interpolate_fft per column
    TRY(dalloc(ctx, &cm->coeffs, N * c));
    Xform I;
    I.logn = ilog2(N);
    I.inverse = true;
    TRY(run_ntt(ctx, I, cols, N, false, cm->coeffs, N, false, c));
    // compute_lde_trace_evaluations: evaluate_offset_fft(blowup, Some(N), h) per column
    TRY(dalloc(ctx, &cm->lde, M * c));
    TRY(evaluate_cosets(ctx, cm->coeffs, N, false, ilog2(N), (unsigned)blowup, H::from_u64(coset_offset), cm->lde, M, false, c));
    if (with_tree) {
        // batch_commit over the rows of the LDE table
        TRY(dalloc(ctx, &cm->nodes, 4 * (2 * M - 1)));
        TRY(build_tree(ctx, cm->lde, M, c, M, cm->nodes));
        TRY(fetch_root(ctx, cm->nodes, root));
    } else {
        CU(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return S252_OK;
}
// Column groups of the upload -> transform pipeline of a host-resident table.  PCIe is about as fast as the transforms
// consume columns, so the pipeline is upload-bound and what stays exposed is the upload of the FIRST group and the
// transforms of the LAST one.  Column-major tables (contiguous DMA at any width) get short groups at both ends
// (weights 1,2,4,..,4,2,1: main commit of the Cairo prover 21.0 -> 19.0 ms).  Row-major tables travel as strided 2-D DMA,
// which is slow for narrow runs (tools/dma2d_bench.py: 19 / 38 / 47 / 50 GB/s for runs of 32 / 64 / 128 / >= 256 bytes),
// so only the first group is short there (weights 1,2,3,4,4,4: measured best, 40.7 ms per C2 step against 41.6 with
// short groups at both ends).  Returns K+1 boundaries.
static std::vector<unsigned> upload_groups(unsigned c, bool column_major) {
    unsigned K = column_major ? 7 : 6;
    if (const char* e = std::getenv("S252_HOST_GROUPS")) { const int v = std::atoi(e); if (v >= 1 && v <= 64) K = (unsigned)v; }
    K = std::min(K, c);
    std::vector<unsigned> cum(K + 1, 0);
    for (unsigned g = 1; g <= K; ++g) {
        unsigned w;
        if (column_major) { const unsigned m = std::min(std::min(g, K - g + 1), 3u); w = m == 3 ? 4 : m; }     // 1,2,4,..,4,2,1
        else w = std::min(g, 4u);                                                                              // 1,2,3,4,4,..
        cum[g] = cum[g - 1] + w;
    }
    std::vector<unsigned> lo(K + 1, 0);
    for (unsigned g = 1; g <= K; ++g) lo[g] = std::max(lo[g - 1] + 1, (unsigned)(((uint64_t)c * cum[g] + cum[K] / 2) / cum[K]));
    lo[K] = c;
    for (unsigned g = K; g-- > 1;) lo[g] = std::min(lo[g], lo[g + 1] - 1);     // K <= c: every group keeps at least one column
    return lo;
}
// How a host table reaches the device (S252_HOST callers):
//   1 = pinned memory (default): strided 2-D DMA of one column group after the other on the copy stream + tile
//       transpose, while the previous group is transformed;
//   2 = pinned memory, read over PCIe by the transposing kernel itself (rows_lw_to_cols_stream; S252_HOST_UPLOAD=zc).
//       Measured slower than the DMA on B200 (C2 step 48.0 vs 40.8 ms): 32-byte element reads make poor PCIe requests;
//   0 = pageable memory (or S252_HOST_UPLOAD=copy): one copy of the whole table on the compute stream (49.0 ms).
static int host_upload_mode(const void* p, const void** dev_alias) {
    *dev_alias = nullptr;
    const char* e = std::getenv("S252_HOST_UPLOAD");
    if (e && !std::strcmp(e, "copy")) return 0;
    cudaPointerAttributes attr{};
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) { cudaGetLastError(); return 0; }
    if (attr.type != cudaMemoryTypeHost) return 0;
    if (!e || std::strcmp(e, "zc") || !attr.devicePointer) return 1;
    *dev_alias = attr.devicePointer;
    return 2;
}

static int lde_from_cols(s252_ctx* ctx, const fe* cols, size_t N, unsigned c, size_t blowup, uint64_t coset_offset, bool with_tree,
                         s252_commit* cm, uint8_t root[32]);
// interpolate_and_commit from a ROW-major LW table in pinned host memory, pipelined: the upload (+ transposition
// + format change) of column group g+1 runs on the copy stream while group g is interpolated and extended.
static int lde_from_pinned_rows(s252_ctx* ctx, const s252_fe* trace, int mode, const void* dev_alias, size_t N, unsigned c,
                                size_t blowup, uint64_t coset_offset, bool with_tree, bool keep_trace, s252_commit* cm,
                                uint8_t root[32]) {
    const size_t M = N * blowup;
    const std::vector<unsigned> lo = upload_groups(c, false);
    const unsigned K = (unsigned)lo.size() - 1;
    std::vector<cudaEvent_t> ev(K, nullptr);
    cudaEvent_t start = nullptr;
    int rc = [&]() -> int {
        Tmp<fe> staged(ctx), cols(ctx);
        if (mode == 1) TRY(dalloc(ctx, &staged.p, N * c));
        TRY(dalloc(ctx, &cols.p, N * c));
        TRY(dalloc(ctx, &cm->coeffs, N * c));
        TRY(dalloc(ctx, &cm->lde, M * c));
        // the blocks may be recycled ones: the copy stream must not overtake work already queued on the compute stream
        CU(ctx, cudaEventCreateWithFlags(&start, cudaEventDisableTiming));
        CU(ctx, cudaEventRecord(start, ctx->stream));
        CU(ctx, cudaStreamWaitEvent(ctx->copy_stream, start, 0));
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
        for (unsigned g = 0; g < K; ++g) {
            const unsigned cg = lo[g + 1] - lo[g];
            const size_t off = (size_t)lo[g] * N;
            if (mode == 2) {
                const fe* src = reinterpret_cast<const fe*>(dev_alias) + lo[g];
                s252::rows_lw_to_cols_stream<<<(unsigned)sms / 2, 256, 0, ctx->copy_stream>>>(src, N, c, cg, cols.p + off, N);
                ctx->launches++;
                CU(ctx, cudaGetLastError());
            } else {
                CU(ctx, cudaMemcpy2DAsync(staged.p + off, (size_t)cg * sizeof(fe), reinterpret_cast<const fe*>(trace) + lo[g],
                                          (size_t)c * sizeof(fe), (size_t)cg * sizeof(fe), N, cudaMemcpyHostToDevice, ctx->copy_stream));
            }
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
            if (mode == 1) {
                prof_begin(ctx, "rows_lw_to_cols");
                prof_work(ctx, 64.0 * N * cg, 0, 0);
                s252::rows_lw_to_cols<<<dim3((unsigned)((N + 31) / 32), (cg + 31) / 32), 256, 0, ctx->stream>>>(staged.p + off, N, cg, cols.p + off, N);
                LAUNCH_CHECK(ctx);
            }
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
    return rc;
}

static int interpolate_lde_impl(s252_ctx* ctx, const s252_fe* trace, size_t n_rows, size_t n_cols, size_t blowup,
                                uint64_t coset_offset, int mem, bool with_tree, s252_commit** out, uint8_t root[32],
                                bool keep_trace = false) {
    if (!ctx || !trace || !out || (with_tree && !root)) return S252_ERR_INVALID;
    *out = nullptr;
    CU(ctx, cudaSetDevice(ctx->device));
    if (!is_pow2(n_rows) || n_cols == 0) FAIL(ctx, S252_ERR_INVALID, "FFTError: trace length %zu is not a power of two", n_rows);
    if (!is_pow2(blowup) || blowup > MAX_COSETS) FAIL(ctx, S252_ERR_INVALID, "blowup factor %zu must be a power of two <= %u", blowup, MAX_COSETS);
    if (coset_offset == 0) FAIL(ctx, S252_ERR_INVALID, "coset offset must be non-zero");
    const size_t N = n_rows, M = n_rows * blowup;
    const unsigned c = (unsigned)n_cols;
    s252_commit* cm = new s252_commit();
    cm->ctx = ctx; cm->n_cols = n_cols; cm->n_rows = M; cm->n_coeffs = N;
    const void* dev_alias = nullptr;
    const int upload = mem == S252_HOST ? host_upload_mode(trace, &dev_alias) : 0;
    int rc = upload ? lde_from_pinned_rows(ctx, trace, upload, dev_alias, N, c, blowup, coset_offset, with_tree, keep_trace, cm, root)
                    : [&]() -> int {
        Tmp<fe> staged(ctx), cols(ctx);
        const fe* dtrace;
        TRY(stage_in(ctx, trace, N * c, mem, staged, &dtrace));
        // TraceTable::cols(): row-major LW -> column-major internal
        TRY(dalloc(ctx, &cols.p, N * c));
        prof_begin(ctx, "rows_lw_to_cols");
        prof_work(ctx, 64.0 * N * c, 0, 0);
        if (c >= 16) {
            s252::rows_lw_to_cols<<<dim3((unsigned)((N + 31) / 32), (c + 31) / 32), 256, 0, ctx->stream>>>(dtrace, N, c, cols.p, N);
        } else {
            // a narrow table (one rank's columns of a sharded trace) leaves most of a 32 x 32 tile empty: one thread per
            // element instead -- coalesced reads along the rows, 32-byte sector writes down the columns
            int sms = 148;
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
            s252::rows_lw_to_cols_stream<<<(unsigned)sms * 8, 256, 0, ctx->stream>>>(dtrace, N, c, c, cols.p, N);
        }
        LAUNCH_CHECK(ctx);
        TRY(lde_from_cols(ctx, cols.p, N, c, blowup, coset_offset, with_tree, cm, root));
        if (keep_trace) { cm->trace = cols.p; cols.p = nullptr; }
        return S252_OK;
    }();
    if (rc != S252_OK) { commit_free(cm); return rc; }
    *out = cm;
    return S252_OK;
}
extern "C" int s252_interpolate_and_commit(s252_ctx* ctx, const s252_fe* trace, size_t n_rows, size_t n_cols, size_t blowup,
                                           uint64_t coset_offset, int mem, s252_commit** out, uint8_t root[32]) {
    NVTX_RANGE("s252_interpolate_and_commit");
    return interpolate_lde_impl(ctx, trace, n_rows, n_cols, blowup, coset_offset, mem, true, out, root);
}
extern "C" int s252_interpolate_and_lde(s252_ctx* ctx, const s252_fe* trace, size_t n_rows, size_t n_cols, size_t blowup,
                                        uint64_t coset_offset, int mem, s252_commit** out) {
    NVTX_RANGE("s252_interpolate_and_lde");
    return interpolate_lde_impl(ctx, trace, n_rows, n_cols, blowup, coset_offset, mem, false, out, nullptr);
}
// batch_commit over column-major columns that are already on this device in the library's internal
// element format (e.g. a row block assembled from the LDE shards of several GPUs).  The columns
// are copied into the handle.
extern "C" int s252_commit_device_columns(s252_ctx* ctx, const void* cols, size_t col_stride, size_t n_cols, size_t n_rows,
                                          s252_commit** out, uint8_t root[32]) {
    NVTX_RANGE("s252_commit_device_columns");
    if (!ctx || !cols || !out || !root) return S252_ERR_INVALID;
    *out = nullptr;
    CU(ctx, cudaSetDevice(ctx->device));
    if (!is_pow2(n_rows) || n_cols == 0 || col_stride < n_rows)
        FAIL(ctx, S252_ERR_INVALID, "merkle tree needs a power-of-two number of leaves (got %zu)", n_rows);
    s252_commit* cm = new s252_commit();
    cm->ctx = ctx; cm->n_cols = n_cols; cm->n_rows = n_rows; cm->n_coeffs = 0;
    int rc = [&]() -> int {
        TRY(dalloc(ctx, &cm->lde, n_rows * n_cols));
        CU(ctx, cudaMemcpy2DAsync(cm->lde, n_rows * sizeof(fe), cols, col_stride * sizeof(fe), n_rows * sizeof(fe), n_cols,
                                  cudaMemcpyDeviceToDevice, ctx->stream));
        TRY(dalloc(ctx, &cm->nodes, 4 * (2 * n_rows - 1)));
        TRY(build_tree(ctx, cm->lde, n_rows, (unsigned)n_cols, n_rows, cm->nodes));
        TRY(fetch_root(ctx, cm->nodes, root));
        return S252_OK;
    }();
    if (rc != S252_OK) { commit_free(cm); return rc; }
    *out = cm;
    return S252_OK;
}

// The same without the copy: the tree is built over the caller's columns, which must stay alive and
// unchanged for as long as the handle is used (a multi-GB row block assembled by an all-to-all).
extern "C" int s252_commit_device_columns_inplace(s252_ctx* ctx, const void* cols, size_t col_stride, size_t n_cols, size_t n_rows,
                                                  s252_commit** out, uint8_t root[32]) {
    NVTX_RANGE("s252_commit_device_columns_inplace");
    if (!ctx || !cols || !out) return S252_ERR_INVALID;          // root == NULL: leave the root on the device (no read-back, no wait)
    *out = nullptr;
    CU(ctx, cudaSetDevice(ctx->device));
    if (!is_pow2(n_rows) || n_cols == 0 || col_stride != n_rows)
        FAIL(ctx, S252_ERR_INVALID, "in-place commit needs a power-of-two number of rows and col_stride == n_rows (got %zu, %zu)", n_rows, col_stride);
    s252_commit* cm = new s252_commit();
    cm->ctx = ctx; cm->n_cols = n_cols; cm->n_rows = n_rows; cm->n_coeffs = 0;
    cm->lde = const_cast<fe*>(reinterpret_cast<const fe*>(cols));
    cm->owns_lde = false;
    int rc = [&]() -> int {
        TRY(dalloc(ctx, &cm->nodes, 4 * (2 * n_rows - 1)));
        TRY(build_tree(ctx, cm->lde, n_rows, (unsigned)n_cols, n_rows, cm->nodes));
        if (root) TRY(fetch_root(ctx, cm->nodes, root));
        return S252_OK;
    }();
    if (rc != S252_OK) { commit_free(cm); return rc; }
    *out = cm;
    return S252_OK;
}

extern "C" int s252_lde_and_commit(s252_ctx* ctx, const s252_fe* polys, size_t n_coeffs, size_t n_polys, size_t domain_size,
                                   size_t blowup, uint64_t coset_offset, int mem, s252_commit** out, uint8_t root[32]) {
    NVTX_RANGE("s252_lde_and_commit");
    if (!ctx || !polys || !out || !root) return S252_ERR_INVALID;
    *out = nullptr;
    CU(ctx, cudaSetDevice(ctx->device));
    if (!is_pow2(domain_size) || !is_pow2(blowup) || blowup > MAX_COSETS || n_polys == 0)
        FAIL(ctx, S252_ERR_INVALID, "domain size and blowup must be powers of two");
    if (n_coeffs > domain_size)
        FAIL(ctx, S252_ERR_INVALID, "polynomials with %zu coefficients exceed the trace domain %zu; use "
                                    "s252_evaluate_polynomial_on_lde_domain + s252_merkle_build", n_coeffs, domain_size);
    if (coset_offset == 0) FAIL(ctx, S252_ERR_INVALID, "coset offset must be non-zero");
    const size_t N = domain_size, M = domain_size * blowup;
    const unsigned c = (unsigned)n_polys;
    s252_commit* cm = new s252_commit();
    cm->ctx = ctx; cm->n_cols = n_polys; cm->n_rows = M; cm->n_coeffs = N;
    int rc = [&]() -> int {
        Tmp<fe> staged(ctx);
        const fe* dpolys;
        TRY(stage_in(ctx, polys, n_coeffs * c, mem, staged, &dpolys));
        TRY(dalloc(ctx, &cm->coeffs, N * c));
        CU(ctx, cudaMemsetAsync(cm->coeffs, 0, N * c * sizeof(fe), ctx->stream));
        for (unsigned j = 0; j < c; ++j)
            if (n_coeffs) TRY(convert_lw_to_internal(ctx, dpolys + j * n_coeffs, cm->coeffs + j * N, n_coeffs));
        TRY(dalloc(ctx, &cm->lde, M * c));
        TRY(evaluate_cosets(ctx, cm->coeffs, N, false, ilog2(N), (unsigned)blowup, H::from_u64(coset_offset), cm->lde, M, false, c));
        TRY(dalloc(ctx, &cm->nodes, 4 * (2 * M - 1)));
        TRY(build_tree(ctx, cm->lde, M, c, M, cm->nodes));
        TRY(fetch_root(ctx, cm->nodes, root));
        return S252_OK;
    }();
    if (rc != S252_OK) { commit_free(cm); return rc; }
    *out = cm;
    return S252_OK;
}

extern "C" int s252_merkle_build(s252_ctx* ctx, const s252_fe* rows, size_t n_rows, size_t n_cols, int mem, s252_commit** out,
                                 uint8_t root[32]) {
    NVTX_RANGE("s252_merkle_build");
    if (!ctx || !rows || !out || !root) return S252_ERR_INVALID;
    *out = nullptr;
    CU(ctx, cudaSetDevice(ctx->device));
    if (!is_pow2(n_rows) || n_cols == 0) FAIL(ctx, S252_ERR_INVALID, "merkle tree needs a power-of-two number of leaves (got %zu)", n_rows);
    const unsigned c = (unsigned)n_cols;
    s252_commit* cm = new s252_commit();
    cm->ctx = ctx; cm->n_cols = n_cols; cm->n_rows = n_rows; cm->n_coeffs = 0;
    int rc = [&]() -> int {
        Tmp<fe> staged(ctx);
        const fe* drows;
        TRY(stage_in(ctx, rows, n_rows * c, mem, staged, &drows));
        TRY(dalloc(ctx, &cm->lde, n_rows * c));
        prof_begin(ctx, "rows_lw_to_cols");
        s252::rows_lw_to_cols<<<dim3((unsigned)((n_rows + 31) / 32), (c + 31) / 32), 256, 0, ctx->stream>>>(drows, n_rows, c, cm->lde, n_rows);
        LAUNCH_CHECK(ctx);
        TRY(dalloc(ctx, &cm->nodes, 4 * (2 * n_rows - 1)));
        TRY(build_tree(ctx, cm->lde, n_rows, c, n_rows, cm->nodes));
        TRY(fetch_root(ctx, cm->nodes, root));
        return S252_OK;
    }();
    if (rc != S252_OK) { commit_free(cm); return rc; }
    *out = cm;
    return S252_OK;
}
extern "C" void s252_commit_destroy(s252_commit* c) {
    if (!c) return;
    cudaSetDevice(c->ctx->device);
    commit_free(c);
}
extern "C" size_t s252_commit_n_cols(const s252_commit* c) { return c->n_cols; }
extern "C" size_t s252_commit_n_rows(const s252_commit* c) { return c->n_rows; }
extern "C" size_t s252_commit_n_coeffs(const s252_commit* c) { return c->n_coeffs; }
extern "C" int s252_commit_root(const s252_commit* c, uint8_t root[32]) {
    s252_ctx* ctx = c->ctx;
    CU(ctx, cudaSetDevice(ctx->device));
    if (!c->nodes) FAIL(ctx, S252_ERR_INVALID, "this handle has no Merkle tree");
    return fetch_root(ctx, c->nodes, root);
}
static int read_internal_as_lw(s252_ctx* ctx, const fe* src, size_t count, s252_fe* out) {
    Tmp<fe> tmp(ctx);
    TRY(dalloc(ctx, &tmp.p, count));
    prof_begin(ctx, "internal_to_lw");
    s252::internal_to_lw<<<(unsigned)((count + 255) / 256), 256, 0, ctx->stream>>>(src, tmp.p, count);
    LAUNCH_CHECK(ctx);
    CU(ctx, cudaMemcpyAsync(out, tmp.p, count * sizeof(fe), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return S252_OK;
}
extern "C" int s252_commit_read_lde(s252_commit* c, size_t col, size_t first, size_t count, s252_fe* out) {
    s252_ctx* ctx = c->ctx;
    CU(ctx, cudaSetDevice(ctx->device));
    if (col >= c->n_cols || first + count > c->n_rows) FAIL(ctx, S252_ERR_RANGE, "LDE read out of range");
    if (count == 0) return S252_OK;
    return read_internal_as_lw(ctx, c->lde + col * c->n_rows + first, count, out);
}
extern "C" int s252_commit_read_coeffs(s252_commit* c, size_t col, s252_fe* out) {
    s252_ctx* ctx = c->ctx;
    CU(ctx, cudaSetDevice(ctx->device));
    if (!c->coeffs || col >= c->n_cols) FAIL(ctx, S252_ERR_RANGE, "coefficient read out of range");
    return read_internal_as_lw(ctx, c->coeffs + col * c->n_coeffs, c->n_coeffs, out);
}
extern "C" int s252_commit_read_nodes(s252_commit* c, size_t first, size_t count, uint8_t* out) {
    s252_ctx* ctx = c->ctx;
    CU(ctx, cudaSetDevice(ctx->device));
    if (!c->nodes || first + count > 2 * c->n_rows - 1) FAIL(ctx, S252_ERR_RANGE, "node read out of range");
    CU(ctx, cudaMemcpyAsync(out, c->nodes + 4 * first, count * 32, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return S252_OK;
}
static int open_common(s252_ctx* ctx, const fe* cols, size_t col_stride, unsigned ncols, const uint64_t* nodes, size_t n_rows,
                       const uint64_t* indices, size_t n_idx, s252_fe* rows_out, uint8_t* paths_out) {
    if (n_idx == 0) return S252_OK;
    for (size_t q = 0; q < n_idx; ++q)
        if (indices[q] >= n_rows) FAIL(ctx, S252_ERR_RANGE, "position %llu is outside the tree (%zu leaves)", (unsigned long long)indices[q], n_rows);
    const unsigned depth = ilog2(n_rows);
    Tmp<unsigned long long> didx(ctx);
    Tmp<fe> drows(ctx);
    Tmp<uint64_t> dpaths(ctx);
    TRY(dalloc(ctx, &didx.p, n_idx));
    CU(ctx, cudaMemcpyAsync(didx.p, indices, n_idx * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (rows_out) {
        TRY(dalloc(ctx, &drows.p, n_idx * ncols));
        prof_begin(ctx, "gather_rows");
        s252::gather_rows<<<(unsigned)((n_idx * ncols + 127) / 128), 128, 0, ctx->stream>>>(cols, col_stride, ncols, didx.p, (unsigned)n_idx, drows.p);
        LAUNCH_CHECK(ctx);
        CU(ctx, cudaMemcpyAsync(rows_out, drows.p, n_idx * ncols * sizeof(fe), cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (paths_out && depth) {
        TRY(dalloc(ctx, &dpaths.p, n_idx * depth * 4));
        prof_begin(ctx, "gather_paths");
        s252::gather_paths<<<(unsigned)((n_idx * depth + 127) / 128), 128, 0, ctx->stream>>>(nodes, depth, didx.p, (unsigned)n_idx, dpaths.p);
        LAUNCH_CHECK(ctx);
        CU(ctx, cudaMemcpyAsync(paths_out, dpaths.p, n_idx * depth * 32, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return S252_OK;
}
extern "C" int s252_commit_open(s252_commit* c, const uint64_t* indices, size_t n_idx, s252_fe* rows_out, uint8_t* paths_out) {
    NVTX_RANGE("s252_commit_open");
    s252_ctx* ctx = c->ctx;
    CU(ctx, cudaSetDevice(ctx->device));
    if (!indices && n_idx) return S252_ERR_INVALID;
    if (paths_out && !c->nodes) FAIL(ctx, S252_ERR_INVALID, "this handle has no Merkle tree");
    return open_common(ctx, c->lde, c->n_rows, (unsigned)c->n_cols, c->nodes, c->n_rows, indices, n_idx, rows_out, paths_out);
}
extern "C" const void* s252_commit_device_lde(const s252_commit* c) { return c->lde; }
extern "C" const void* s252_commit_device_coeffs(const s252_commit* c) { return c->coeffs; }
extern "C" const void* s252_commit_device_nodes(const s252_commit* c) { return c->nodes; }

// --------------------------------------------------------------------------------------------
// FRI
static void fri_free(s252_fri* f) {
    if (!f) return;
    for (auto& l : f->layers) { dfree(f->ctx, l.evals); dfree(f->ctx, l.nodes); }
    delete f;
}
// FRI commit phase once layer 0 (f->layers[0].evals, domain_size evaluations on the coset h*<w>) is
// resident: trees, transcript, folds, last value (fri/mod.rs:33-69).
//
// The whole phase is queued without a host round trip: the DefaultTranscript sponge continues on the device
// (s252::FriChain), every tree's last block appends its root and samples the next zeta, layers of at most
// 2^FRI_TAIL_LOG_MAX evaluations are finished by one single-block kernel, and the host reads all roots and the
// last value back ONCE and replays the same appends on its own transcript (checking the last zeta).
static unsigned fri_tail_log_max() {
    if (const char* e = std::getenv("S252_FRI_TAIL_LOG")) { const int v = std::atoi(e); if (v >= 0 && v <= (int)s252::FRI_TAIL_LOG_MAX) return (unsigned)v; }
    return 11;      // measured on the C2 step (B200): 0 -> 32.30, 11 -> 32.23, 13 -> 32.61, 14 -> 33.15 ms
}
static int fri_from_layer0(s252_ctx* ctx, s252_fri* f, size_t number_layers, s252_transcript* transcript, fe h,
                           size_t domain_size, s252_fe* last_value, uint8_t* roots_out) {
    const unsigned logM = ilog2(domain_size);
    fe wM, w_inv;
    H::primitive_root(logM, &wM);
    w_inv = H::inv(wM);
    const fe* inv_tw = nullptr;
    if (domain_size >= 2) TRY(get_power_table(ctx, domain_size / 2, w_inv, H::one(), &inv_tw));
    const fe inv2 = H::inv(H::from_u64(2));
    const size_t L = number_layers;
    if (L > (size_t)s252::FRI_MAX_LAYERS) FAIL(ctx, S252_ERR_INVALID, "at most %d FRI layers are supported", s252::FRI_MAX_LAYERS);
    if ((domain_size >> (L ? L : 1)) == 0) FAIL(ctx, S252_ERR_INVALID, "FRI layer of size %zu cannot be folded", domain_size >> (L ? L - 1 : 0));
    fe lv;
    if (L == 0) {
        // no layer is committed (fri/mod.rs:58-69 only): zeta comes straight from the host transcript
        const fe zeta = transcript->to_field();
        const fe cfac = H::mul(zeta, H::mul(inv2, H::inv(h)));
        const size_t half = domain_size / 2;
        Tmp<fe> rem(ctx);
        TRY(dalloc(ctx, &rem.p, half));
        prof_begin(ctx, "fri_fold_commit");
        prof_work(ctx, 32.0 * domain_size + 32.0 * half, 3.2 * half, 0);
        s252::fri_fold_commit<<<(unsigned)((half + 127) / 128), 128, 0, ctx->stream>>>(f->layers[0].evals, half, inv_tw, 1ull, cfac, nullptr,
                                                                                      inv2, rem.p, nullptr);
        LAUNCH_CHECK(ctx);
        std::vector<fe> tail(half);
        CU(ctx, cudaMemcpyAsync(tail.data(), rem.p, half * sizeof(fe), cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        fe acc = H::zero();
        for (size_t i = 0; i < half; ++i) acc = H::add(acc, tail[i]);
        lv = H::mul(acc, H::inv(H::from_u64((uint64_t)half)));
        dfree(ctx, f->layers.back().evals);
        f->layers.pop_back();
    } else {
        // ---- the chain: the transcript as it stands + 1/(2 h^(2^k)) for every fold
        std::vector<s252::FriChain> hc(1);
        std::memset(hc.data(), 0, sizeof(s252::FriChain));
        std::memcpy(hc[0].sponge, transcript->k.lanes(), 200);
        hc[0].fill = (unsigned)transcript->k.fill();
        {
            fe hk = h;
            for (size_t k = 0; k < L; ++k) { hc[0].inv2h[k] = H::mul(inv2, H::inv(hk)); hk = H::sqr(hk); }
        }
        Tmp<s252::FriChain> chain(ctx);
        TRY(dalloc(ctx, &chain.p, 1));
        CU(ctx, cudaMemcpyAsync(chain.p, hc.data(), sizeof(s252::FriChain), cudaMemcpyHostToDevice, ctx->stream));
        // ---- layer 0 (fri/mod.rs:33-37)
        TRY(dalloc(ctx, &f->layers[0].nodes, 4 * (2 * domain_size - 1)));
        TRY(build_tree(ctx, f->layers[0].evals, domain_size, 1, domain_size, f->layers[0].nodes, chain.p));
        // ---- folds k = 1..L: fold k turns layer k-1 (size M >> (k-1)) into layer k; committed iff k < L
        const unsigned tail_log = fri_tail_log_max();
        size_t k = 1;
        for (; k <= L && ilog2(domain_size >> (k - 1)) > tail_log; ++k) {
            const size_t size = domain_size >> (k - 1), half = size / 2;
            const bool commit = k < L;
            FriLayerDev nxt;
            nxt.size = half;
            TRY(dalloc(ctx, &nxt.evals, half));
            f->layers.push_back(nxt);
            FriLayerDev& Lk = f->layers.back();
            if (commit) TRY(dalloc(ctx, &Lk.nodes, 4 * (2 * half - 1)));
            prof_begin(ctx, "fri_fold_commit");
            prof_work(ctx, 32.0 * size + 32.0 * half + (commit ? 32.0 * half : 0.0), 3.2 * half, commit ? (double)half : 0.0);
            s252::fri_fold_commit<<<(unsigned)((half + 127) / 128), 128, 0, ctx->stream>>>(
                f->layers[k - 1].evals, half, inv_tw, (unsigned long long)(domain_size / size), fe{}, chain.p, inv2, Lk.evals,
                commit ? Lk.nodes + 4 * (half - 1) : nullptr);
            LAUNCH_CHECK(ctx);
            if (commit) TRY(build_tree_nodes(ctx, half, Lk.nodes, chain.p));
        }
        const bool tail = k <= L;
        if (tail) {
            s252::FriTail T{};
            T.chain = chain.p;
            T.in = f->layers[k - 1].evals;
            T.log_size = ilog2(domain_size >> (k - 1));
            T.n_commit = (unsigned)(L - k);
            T.inv_tw = inv_tw;
            T.tw_stride = (unsigned long long)(domain_size / (domain_size >> (k - 1)));
            T.inv2 = inv2;
            T.inv_last = H::inv(H::from_u64((uint64_t)(domain_size >> L)));
            double perms = 0, muls = 0;
            for (size_t j = k; j <= L; ++j) {
                const size_t half = domain_size >> j;
                FriLayerDev nxt;
                nxt.size = half;
                TRY(dalloc(ctx, &nxt.evals, half));
                f->layers.push_back(nxt);
                if (j < L) TRY(dalloc(ctx, &f->layers.back().nodes, 4 * (2 * half - 1)));
                T.evals[j - k] = f->layers.back().evals;
                T.nodes[j - k] = f->layers.back().nodes;
                muls += 3.2 * half;
                if (j < L) perms += 2.0 * half;
            }
            prof_begin(ctx, "fri_tail_kernel");
            prof_work(ctx, 32.0 * 3 * (domain_size >> (k - 1)), muls, perms);
            s252::fri_tail_kernel<<<1, s252::FRI_TAIL_THREADS, 0, ctx->stream>>>(T);
            LAUNCH_CHECK(ctx);
        }
        // ---- one read-back: roots, last zeta, last value
        CU(ctx, cudaMemcpyAsync(hc.data(), chain.p, sizeof(s252::FriChain), cudaMemcpyDeviceToHost, ctx->stream));
        std::vector<fe> rem;
        const size_t left = domain_size >> L;
        if (!tail) {
            rem.resize(left);
            CU(ctx, cudaMemcpyAsync(rem.data(), f->layers.back().evals, left * sizeof(fe), cudaMemcpyDeviceToHost, ctx->stream));
        }
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        // ---- replay on the host transcript (fri/mod.rs:37,41,54,58)
        fe zeta = H::zero();
        for (size_t j = 0; j < L; ++j) {
            if (j > 0) zeta = transcript->to_field();
            transcript->append(reinterpret_cast<const uint8_t*>(hc[0].roots[j]), 32);
            if (roots_out) std::memcpy(roots_out + 32 * j, hc[0].roots[j], 32);
        }
        zeta = transcript->to_field();
        if (hc[0].layer != L || !H::eq(zeta, hc[0].zeta)) FAIL(ctx, S252_ERR_CUDA, "internal: the device-side transcript diverged from the host's");
        if (tail) {
            lv = hc[0].last_value;
        } else {
            // last_value = coefficient 0 of the fully folded polynomial = mean of its evaluations on the remaining
            // coset (it has at most `left` coefficients because p0 has at most domain_size)
            fe acc = H::zero();
            for (size_t i = 0; i < left; ++i) acc = H::add(acc, rem[i]);
            lv = H::mul(acc, H::inv(H::from_u64((uint64_t)left)));
        }
        // the folded remainder is not a FriLayer of the reference: drop it
        dfree(ctx, f->layers.back().evals);
        f->layers.pop_back();
    }
    H::to_lw(lv, last_value->limbs);
    uint8_t be[32];
    H::to_bytes_be(lv, be);
    transcript->append(be, 32);                             // fri/mod.rs:69
    return S252_OK;
}

extern "C" int s252_fri_commit_phase(s252_ctx* ctx, size_t number_layers, const s252_fe* p0, size_t n_coeffs,
                                     s252_transcript* transcript, const s252_fe* coset_offset, size_t domain_size, int mem,
                                     s252_fri** out, s252_fe* last_value, uint8_t* roots_out) {
    NVTX_RANGE("s252_fri_commit_phase");
    if (!ctx || !transcript || !coset_offset || !out || !last_value || (!p0 && n_coeffs)) return S252_ERR_INVALID;
    *out = nullptr;
    CU(ctx, cudaSetDevice(ctx->device));
    if (!is_pow2(domain_size)) FAIL(ctx, S252_ERR_INVALID, "FRI domain size %zu is not a power of two", domain_size);
    if (n_coeffs > domain_size) FAIL(ctx, S252_ERR_INVALID, "p0 has %zu coefficients, more than the domain size %zu", n_coeffs, domain_size);
    if (number_layers > ilog2(domain_size)) FAIL(ctx, S252_ERR_INVALID, "%zu FRI layers do not fit a domain of size %zu", number_layers, domain_size);
    const fe h = H::from_lw(coset_offset->limbs);
    if (H::is_zero(h)) FAIL(ctx, S252_ERR_INVALID, "coset offset must be non-zero");
    s252_fri* f = new s252_fri();
    f->ctx = ctx; f->domain_size = domain_size;
    int rc = [&]() -> int {
        Tmp<fe> staged(ctx);
        const fe* dp0 = nullptr;
        if (n_coeffs) TRY(stage_in(ctx, p0, n_coeffs, mem, staged, &dp0));
        // layer 0: FriLayer::new(p0, h, domain_size)
        FriLayerDev cur;
        cur.size = domain_size;
        TRY(dalloc(ctx, &cur.evals, domain_size));
        f->layers.push_back(cur);   // owned by f from here on
        TRY(evaluate_from_lw(ctx, dp0, n_coeffs, domain_size, h, f->layers[0].evals, false));
        return fri_from_layer0(ctx, f, number_layers, transcript, h, domain_size, last_value, roots_out);
    }();
    if (rc != S252_OK) { fri_free(f); return rc; }
    *out = f;
    return S252_OK;
}

// Frame::get_trace_evaluations (src/starks/frame.rs:67-83) and the H1/H2 evaluations of round 3
// (prover.rs:296-300) from the coefficients resident in a commit handle.
extern "C" int s252_commit_evaluate_at(s252_commit* c, const s252_fe* points, size_t n_points, s252_fe* out, size_t out_stride,
                                       size_t col_offset) {
    NVTX_RANGE("s252_commit_evaluate_at");
    if (!c || !points || !out) return S252_ERR_INVALID;
    s252_ctx* ctx = c->ctx;
    CU(ctx, cudaSetDevice(ctx->device));
    if (!c->coeffs) FAIL(ctx, S252_ERR_INVALID, "this handle keeps no coefficients");
    if (col_offset + c->n_cols > out_stride) FAIL(ctx, S252_ERR_INVALID, "output row of %zu elements cannot hold columns %zu..%zu", out_stride, col_offset, col_offset + c->n_cols);
    if (n_points == 0) return S252_OK;
    if (n_points > (size_t)s252::EVAL_MAX_POINTS) {
        // more points than one launch handles: split
        for (size_t p0 = 0; p0 < n_points; p0 += s252::EVAL_MAX_POINTS) {
            const size_t cnt = std::min<size_t>(s252::EVAL_MAX_POINTS, n_points - p0);
            TRY(s252_commit_evaluate_at(c, points + p0, cnt, out + p0 * out_stride, out_stride, col_offset));
        }
        return S252_OK;
    }
    // splits: enough blocks to fill the GPU, chains no shorter than ~32 coefficients
    unsigned splits = 1;
    while (splits < 64 && (size_t)c->n_cols * splits < 1024 && c->n_coeffs / ((size_t)s252::EVAL_THREADS * splits * 2) >= 32) splits *= 2;
    Tmp<fe> dx(ctx), dxi(ctx), dpart(ctx);
    const size_t nparts = c->n_cols * splits * n_points;
    TRY(dalloc(ctx, &dx.p, n_points));
    TRY(dalloc(ctx, &dxi.p, n_points));
    TRY(dalloc(ctx, &dpart.p, nparts));
    CU(ctx, cudaMemcpyAsync(dx.p, points, n_points * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream));
    TRY(convert_lw_to_internal(ctx, dx.p, dxi.p, n_points));
    prof_begin(ctx, "poly_eval_points");
    prof_work(ctx, 32.0 * c->n_coeffs * c->n_cols, (double)c->n_coeffs * c->n_cols * n_points, 0);
    const dim3 grid((unsigned)c->n_cols, splits);
    switch (n_points) {
        case 1: s252::poly_eval_points<1><<<grid, s252::EVAL_THREADS, 0, ctx->stream>>>(c->coeffs, c->n_coeffs, c->n_coeffs, dxi.p, dpart.p, splits); break;
        case 2: s252::poly_eval_points<2><<<grid, s252::EVAL_THREADS, 0, ctx->stream>>>(c->coeffs, c->n_coeffs, c->n_coeffs, dxi.p, dpart.p, splits); break;
        case 3: s252::poly_eval_points<3><<<grid, s252::EVAL_THREADS, 0, ctx->stream>>>(c->coeffs, c->n_coeffs, c->n_coeffs, dxi.p, dpart.p, splits); break;
        default: s252::poly_eval_points<4><<<grid, s252::EVAL_THREADS, 0, ctx->stream>>>(c->coeffs, c->n_coeffs, c->n_coeffs, dxi.p, dpart.p, splits); break;
    }
    LAUNCH_CHECK(ctx);
    std::vector<fe> parts(nparts);
    CU(ctx, cudaMemcpyAsync(parts.data(), dpart.p, nparts * sizeof(fe), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    for (size_t j = 0; j < c->n_cols; ++j)
        for (size_t p = 0; p < n_points; ++p) {
            fe acc = H::zero();
            for (unsigned s = 0; s < splits; ++s) acc = H::add(acc, parts[(j * splits + s) * n_points + p]);
            H::to_lw(acc, out[p * out_stride + col_offset + j].limbs);
        }
    return S252_OK;
}

// multiplicative scans (kernels in cairo.cuh): inclusive product of data[0..n) in place (suffix products when reverse)
static int scan_mul(s252_ctx* ctx, fe* data, size_t n, bool reverse) {
    const size_t tiles = (n + s252::SCAN_TILE - 1) / s252::SCAN_TILE;
    if (tiles <= 1) {
        prof_begin(ctx, "scan_mul_tiles");
        s252::scan_mul_tiles<<<1, s252::SCAN_THREADS, 0, ctx->stream>>>(data, n, nullptr, reverse);
        LAUNCH_CHECK(ctx);
        return S252_OK;
    }
    Tmp<fe> totals(ctx);
    TRY(dalloc(ctx, &totals.p, tiles));
    prof_begin(ctx, "scan_mul_tiles");
    prof_work(ctx, 64.0 * n, 2.0 * n, 0);
    s252::scan_mul_tiles<<<(unsigned)tiles, s252::SCAN_THREADS, 0, ctx->stream>>>(data, n, totals.p, reverse);
    LAUNCH_CHECK(ctx);
    TRY(scan_mul(ctx, totals.p, tiles, false));
    prof_begin(ctx, "scan_mul_apply");
    prof_work(ctx, 64.0 * n, (double)n, 0);
    s252::scan_mul_apply<<<(unsigned)((n - s252::SCAN_TILE + 255) / 256), 256, 0, ctx->stream>>>(data, n, totals.p, reverse);
    LAUNCH_CHECK(ctx);
    return S252_OK;
}
static int read_fe(s252_ctx* ctx, const fe* dev, fe* out) {
    CU(ctx, cudaMemcpyAsync(out, dev, sizeof(fe), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return S252_OK;
}
// out[i] = 1 / (dom[i] - c) for a whole array with ONE field inversion (on the host): prefix and suffix
// products are two scans, 1/x_i = prefix[i-1] * suffix[i+1] / total.
static int invert_shifted(s252_ctx* ctx, const fe* dom, size_t n, const fe& c, fe* out) {
    Tmp<fe> a(ctx), b(ctx);
    TRY(dalloc(ctx, &a.p, n));
    TRY(dalloc(ctx, &b.p, n));
    prof_begin(ctx, "sub_const2");
    prof_work(ctx, 96.0 * n, 0, 0);
    s252::sub_const2<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(dom, a.p, b.p, n, c);
    LAUNCH_CHECK(ctx);
    TRY(scan_mul(ctx, a.p, n, false));
    TRY(scan_mul(ctx, b.p, n, true));
    fe total;
    TRY(read_fe(ctx, a.p + (n - 1), &total));
    if (H::is_zero(total)) FAIL(ctx, S252_ERR_INVALID, "division by zero: the point lies on the evaluation domain");
    prof_begin(ctx, "batch_inverse_finish");
    prof_work(ctx, 96.0 * n, 2.0 * n, 0);
    s252::batch_inverse_finish<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(out, a.p, b.p, n, H::inv(total));
    LAUNCH_CHECK(ctx);
    return S252_OK;
}

// The DEEP composition polynomial on a block of LDE rows [row0, row0 + rows): tables[t] holds the block's
// rows of table t column-major with stride strides[t]; the last table is (H1, H2).  out: [rows].
struct DeepTables {
    const fe* cols[s252::DEEP_MAX_TABLES];
    size_t strides[s252::DEEP_MAX_TABLES];
    unsigned ncols[s252::DEEP_MAX_TABLES];
    unsigned ntables;
};
static int deep_evaluate_rows(s252_ctx* ctx, const DeepTables& T, size_t row0, size_t rows, size_t M, size_t Ntrace, const s252_fe* z,
                              const uint64_t* transition_offsets, size_t n_offsets, const s252_fe* trace_ood, const s252_fe* h1_z2,
                              const s252_fe* h2_z2, const s252_fe* gamma, const s252_fe* gamma_p, const s252_fe* trace_gammas,
                              uint64_t coset_offset, fe* out) {
    if (T.ntables < 2 || T.ntables > (unsigned)s252::DEEP_MAX_TABLES) FAIL(ctx, S252_ERR_INVALID, "between 1 and %d trace tables are supported", s252::DEEP_MAX_TABLES - 1);
    if (n_offsets == 0 || n_offsets > (size_t)s252::DEEP_MAX_K) FAIL(ctx, S252_ERR_INVALID, "between 1 and %d frame rows are supported", s252::DEEP_MAX_K);
    if (coset_offset == 0) FAIL(ctx, S252_ERR_INVALID, "coset offset must be non-zero");
    if (!is_pow2(M) || Ntrace == 0 || rows == 0 || row0 + rows > M) FAIL(ctx, S252_ERR_INVALID, "bad domain sizes for the DEEP polynomial");
    s252::DeepParams P{};
    size_t total_cols = 0;
    for (unsigned i = 0; i < T.ntables; ++i) {
        P.cols[i] = T.cols[i]; P.strides[i] = T.strides[i]; P.ncols[i] = T.ncols[i];
        if (i + 1 < T.ntables) total_cols += T.ncols[i];
    }
    if (T.ncols[T.ntables - 1] != 2) FAIL(ctx, S252_ERR_INVALID, "the composition table must hold H1 and H2");
    P.ntables = T.ntables;
    P.K = (unsigned)n_offsets;
    P.m = M;
    const unsigned K = (unsigned)n_offsets;
    fe g;
    if (!H::primitive_root(ilog2(Ntrace), &g)) FAIL(ctx, S252_ERR_INVALID, "no trace root of unity");
    const fe zz = H::from_lw(z->limbs);
    const fe gam = H::from_lw(gamma->limbs), gamp = H::from_lw(gamma_p->limbs);
    std::vector<fe> gammas(total_cols * K + 2);
    for (size_t i = 0; i < total_cols * K; ++i) gammas[i] = H::from_lw(trace_gammas[i].limbs);
    gammas[total_cols * K] = gam;
    gammas[total_cols * K + 1] = gamp;
    const size_t blowup = M / Ntrace;
    const fe ginv1 = H::inv(g);
    for (unsigned k = 0; k < K; ++k) {
        fe acc = H::zero();
        for (size_t j = 0; j < total_cols; ++j) acc = H::add(acc, H::mul(gammas[j * K + k], H::from_lw(trace_ood[k * total_cols + j].limbs)));
        P.ck[k] = acc;
        P.rot[k] = (blowup * (transition_offsets[k] % Ntrace)) % M;
        P.ginv[k] = H::pow_u64(ginv1, transition_offsets[k]);
    }
    P.cz2 = H::add(H::mul(gam, H::from_lw(h1_z2->limbs)), H::mul(gamp, H::from_lw(h2_z2->limbs)));
    const fe h = H::from_u64(coset_offset);
    fe wM;
    H::primitive_root(ilog2(M), &wM);
    Tmp<fe> dg(ctx), dU(ctx), dV(ctx);
    TRY(dalloc(ctx, &dg.p, gammas.size()));
    CU(ctx, cudaMemcpyAsync(dg.p, gammas.data(), gammas.size() * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream));
    P.gammas = dg.p;
    // inverse tables over the LDE coset x_i = h w^i; the frame offsets index U by rotation.  A block of rows (one rank of a
    // sharded proof) needs U only on its rows plus the halo the largest rotation reaches back to, and V only on its rows:
    // the scans then run over the block instead of the whole coset.
    const fe* dom;
    TRY(get_power_table(ctx, M, wM, h, &dom));
    unsigned long long maxrot = 0;
    for (unsigned k = 0; k < K; ++k) maxrot = std::max<unsigned long long>(maxrot, P.rot[k]);
    const bool block = rows < M && rows + maxrot < M;
    const size_t u_len = block ? rows + (size_t)maxrot : M, v_len = block ? rows : M;
    P.u_base = block ? (row0 + M - maxrot) % M : 0;
    P.v_base = block ? row0 : 0;
    Tmp<fe> dom_ext(ctx);
    const fe* dom_u = dom + P.u_base;
    if (block && P.u_base + u_len > M) {                        // the halo wraps around the end of the coset (rank 0)
        TRY(dalloc(ctx, &dom_ext.p, u_len));
        const size_t first = M - P.u_base;
        CU(ctx, cudaMemcpyAsync(dom_ext.p, dom + P.u_base, first * sizeof(fe), cudaMemcpyDeviceToDevice, ctx->stream));
        CU(ctx, cudaMemcpyAsync(dom_ext.p + first, dom, (u_len - first) * sizeof(fe), cudaMemcpyDeviceToDevice, ctx->stream));
        dom_u = dom_ext.p;
    }
    TRY(dalloc(ctx, &dU.p, u_len));
    TRY(dalloc(ctx, &dV.p, v_len));
    TRY(invert_shifted(ctx, dom_u, u_len, zz, dU.p));
    TRY(invert_shifted(ctx, dom + P.v_base, v_len, H::sqr(zz), dV.p));
    P.U = dU.p; P.V = dV.p;
    P.out = out;
    P.row0 = row0; P.rows = rows;
    const unsigned blocks = (unsigned)((rows + s252::DEEP_THREADS - 1) / s252::DEEP_THREADS);
    prof_begin(ctx, "deep_composition_kernel");
    prof_work(ctx, 32.0 * rows * (total_cols + 3 + K + 1), (double)rows * (total_cols * K + 2 + 2 * (K + 1)), 0);
#define S252_DEEP_LAUNCH(KK)                                                                   \
    case KK:                                                                                   \
        s252::deep_composition_kernel<KK><<<blocks, s252::DEEP_THREADS, 0, ctx->stream>>>(P);  \
        break;
    switch (K) {
        S252_DEEP_LAUNCH(1)
        S252_DEEP_LAUNCH(2)
        S252_DEEP_LAUNCH(3)
        S252_DEEP_LAUNCH(4)
    }
#undef S252_DEEP_LAUNCH
    LAUNCH_CHECK(ctx);
    CU(ctx, cudaStreamSynchronize(ctx->stream));   // the host vector / device temporaries stay alive until the kernel is done
    return S252_OK;
}

// Round 4 from the resident commits: DEEP composition polynomial as evaluations on the LDE coset
// (replaces compute_deep_composition_poly, prover.rs:410-482, and FRI layer 0's transform), then
// fri_commit_phase (fri/mod.rs:20-72).
extern "C" int s252_fri_commit_phase_deep(s252_ctx* ctx, size_t number_layers, s252_commit* const* trace_commits,
                                          size_t n_trace_commits, s252_commit* composition_commit, const s252_fe* z,
                                          const uint64_t* transition_offsets, size_t n_offsets, const s252_fe* trace_ood,
                                          const s252_fe* h1_z2, const s252_fe* h2_z2, const s252_fe* gamma,
                                          const s252_fe* gamma_p, const s252_fe* trace_gammas, s252_transcript* transcript,
                                          uint64_t coset_offset, s252_fri** out, s252_fe* last_value, uint8_t* roots_out) {
    NVTX_RANGE("s252_fri_commit_phase_deep");
    if (!ctx || !trace_commits || !composition_commit || !z || !transition_offsets || !trace_ood || !h1_z2 || !h2_z2 || !gamma ||
        !gamma_p || !trace_gammas || !transcript || !out || !last_value)
        return S252_ERR_INVALID;
    *out = nullptr;
    CU(ctx, cudaSetDevice(ctx->device));
    if (n_trace_commits == 0 || n_trace_commits + 1 > s252::DEEP_MAX_TABLES) FAIL(ctx, S252_ERR_INVALID, "between 1 and %d trace tables are supported", s252::DEEP_MAX_TABLES - 1);
    if (composition_commit->n_cols != 2) FAIL(ctx, S252_ERR_INVALID, "the composition commit must hold H1 and H2");
    const size_t M = composition_commit->n_rows, Ntrace = trace_commits[0]->n_coeffs;
    if (!is_pow2(M) || Ntrace == 0 || number_layers > ilog2(M)) FAIL(ctx, S252_ERR_INVALID, "bad domain sizes for FRI");
    DeepTables T{};
    for (size_t i = 0; i < n_trace_commits; ++i) {
        const s252_commit* tc = trace_commits[i];
        if (tc->n_rows != M || tc->ctx != ctx) FAIL(ctx, S252_ERR_INVALID, "trace commit %zu does not share the LDE domain / context", i);
        T.cols[i] = tc->lde; T.strides[i] = tc->n_rows; T.ncols[i] = (unsigned)tc->n_cols;
    }
    T.cols[n_trace_commits] = composition_commit->lde;
    T.strides[n_trace_commits] = M;
    T.ncols[n_trace_commits] = 2;
    T.ntables = (unsigned)n_trace_commits + 1;
    s252_fri* f = new s252_fri();
    f->ctx = ctx; f->domain_size = M;
    int rc = [&]() -> int {
        FriLayerDev cur;
        cur.size = M;
        TRY(dalloc(ctx, &cur.evals, M));
        f->layers.push_back(cur);
        TRY(deep_evaluate_rows(ctx, T, 0, M, M, Ntrace, z, transition_offsets, n_offsets, trace_ood, h1_z2, h2_z2, gamma, gamma_p,
                               trace_gammas, coset_offset, f->layers[0].evals));
        return fri_from_layer0(ctx, f, number_layers, transcript, H::from_u64(coset_offset), M, last_value, roots_out);
    }();
    if (rc != S252_OK) { fri_free(f); return rc; }
    *out = f;
    return S252_OK;
}
// fri_commit_phase (fri/mod.rs:20-72) from layer 0 given as EVALUATIONS on the LDE coset, resident on this
// device in the library's internal element format (e.g. the DEEP polynomial gathered from row blocks).
extern "C" int s252_fri_commit_phase_evals(s252_ctx* ctx, size_t number_layers, const void* p0_evals, size_t domain_size,
                                           s252_transcript* transcript, uint64_t coset_offset, s252_fri** out, s252_fe* last_value,
                                           uint8_t* roots_out) {
    NVTX_RANGE("s252_fri_commit_phase_evals");
    if (!ctx || !p0_evals || !transcript || !out || !last_value) return S252_ERR_INVALID;
    *out = nullptr;
    CU(ctx, cudaSetDevice(ctx->device));
    if (!is_pow2(domain_size) || number_layers > ilog2(domain_size) || coset_offset == 0) FAIL(ctx, S252_ERR_INVALID, "bad domain sizes for FRI");
    s252_fri* f = new s252_fri();
    f->ctx = ctx; f->domain_size = domain_size;
    int rc = [&]() -> int {
        FriLayerDev cur;
        cur.size = domain_size;
        TRY(dalloc(ctx, &cur.evals, domain_size));
        f->layers.push_back(cur);
        CU(ctx, cudaMemcpyAsync(f->layers[0].evals, p0_evals, domain_size * sizeof(fe), cudaMemcpyDeviceToDevice, ctx->stream));
        return fri_from_layer0(ctx, f, number_layers, transcript, H::from_u64(coset_offset), domain_size, last_value, roots_out);
    }();
    if (rc != S252_OK) { fri_free(f); return rc; }
    *out = f;
    return S252_OK;
}
// ---- layer-by-layer interface (fri_layer0 / fri_fold_commit) for a caller that keeps its own transcript:
// FriLayer::new(p0, offset, domain_size) (fri_commitment.rs:30-47), then per layer fold_polynomial(zeta) + FriLayer::new
// (fri/mod.rs:43-54), then the last fold (fri/mod.rs:58-66).
extern "C" int s252_fri_layer0(s252_ctx* ctx, const s252_fe* p0, size_t n_coeffs, const s252_fe* coset_offset, size_t domain_size, int mem,
                               s252_fri** out, uint8_t root[32]) {
    NVTX_RANGE("s252_fri_layer0");
    if (!ctx || !coset_offset || !out || !root || (!p0 && n_coeffs)) return S252_ERR_INVALID;
    *out = nullptr;
    CU(ctx, cudaSetDevice(ctx->device));
    if (!is_pow2(domain_size) || n_coeffs > domain_size) FAIL(ctx, S252_ERR_INVALID, "p0 with %zu coefficients does not fit a domain of %zu", n_coeffs, domain_size);
    const fe h = H::from_lw(coset_offset->limbs);
    if (H::is_zero(h)) FAIL(ctx, S252_ERR_INVALID, "coset offset must be non-zero");
    s252_fri* f = new s252_fri();
    f->ctx = ctx; f->domain_size = domain_size; f->h0 = h;
    int rc = [&]() -> int {
        Tmp<fe> staged(ctx);
        const fe* dp0 = nullptr;
        if (n_coeffs) TRY(stage_in(ctx, p0, n_coeffs, mem, staged, &dp0));
        FriLayerDev cur;
        cur.size = domain_size;
        TRY(dalloc(ctx, &cur.evals, domain_size));
        f->layers.push_back(cur);
        TRY(evaluate_from_lw(ctx, dp0, n_coeffs, domain_size, h, f->layers[0].evals, false));
        TRY(dalloc(ctx, &f->layers[0].nodes, 4 * (2 * domain_size - 1)));
        TRY(build_tree(ctx, f->layers[0].evals, domain_size, 1, domain_size, f->layers[0].nodes));
        return fetch_root(ctx, f->layers[0].nodes, root);
    }();
    if (rc != S252_OK) { fri_free(f); return rc; }
    *out = f;
    return S252_OK;
}
static int fri_fold_step(s252_fri* f, const s252_fe* zeta, bool commit, FriLayerDev* nxt) {
    s252_ctx* ctx = f->ctx;
    const size_t k = f->layers.size();                                 // index of the layer being produced
    const size_t size = f->layers.back().size, half = size / 2;
    if (half == 0) FAIL(ctx, S252_ERR_INVALID, "FRI layer of size %zu cannot be folded", size);
    fe w;
    H::primitive_root(ilog2(f->domain_size), &w);
    const fe* inv_tw;
    TRY(get_power_table(ctx, f->domain_size / 2, H::inv(w), H::one(), &inv_tw));
    fe hk = f->h0;
    for (size_t j = 0; j + 1 < k; ++j) hk = H::sqr(hk);                 // offset of layer k-1
    const fe inv2 = H::inv(H::from_u64(2));
    const fe cfac = H::mul(H::from_lw(zeta->limbs), H::mul(inv2, H::inv(hk)));
    nxt->size = half;
    TRY(dalloc(ctx, &nxt->evals, half));
    if (commit) TRY(dalloc(ctx, &nxt->nodes, 4 * (2 * half - 1)));
    prof_begin(ctx, "fri_fold_commit");
    prof_work(ctx, 32.0 * size + 32.0 * half + (commit ? 32.0 * half : 0.0), 3.2 * half, commit ? (double)half : 0.0);
    s252::fri_fold_commit<<<(unsigned)((half + 127) / 128), 128, 0, ctx->stream>>>(f->layers.back().evals, half, inv_tw,
                                                                                  (unsigned long long)(f->domain_size / size), cfac, nullptr, inv2,
                                                                                  nxt->evals, commit ? nxt->nodes + 4 * (half - 1) : nullptr);
    LAUNCH_CHECK(ctx);
    return S252_OK;
}
extern "C" int s252_fri_fold_commit(s252_fri* f, const s252_fe* zeta, uint8_t root[32]) {
    NVTX_RANGE("s252_fri_fold_commit");
    if (!f || !zeta || !root || f->layers.empty()) return S252_ERR_INVALID;
    s252_ctx* ctx = f->ctx;
    CU(ctx, cudaSetDevice(ctx->device));
    FriLayerDev nxt;
    int rc = fri_fold_step(f, zeta, true, &nxt);
    if (rc == S252_OK) rc = build_tree_nodes(ctx, nxt.size, nxt.nodes);
    if (rc == S252_OK) rc = fetch_root(ctx, nxt.nodes, root);
    if (rc != S252_OK) { dfree(ctx, nxt.evals); dfree(ctx, nxt.nodes); return rc; }
    f->layers.push_back(nxt);
    return S252_OK;
}
extern "C" int s252_fri_fold_last(s252_fri* f, const s252_fe* zeta, s252_fe* last_value) {
    NVTX_RANGE("s252_fri_fold_last");
    if (!f || !zeta || !last_value || f->layers.empty()) return S252_ERR_INVALID;
    s252_ctx* ctx = f->ctx;
    CU(ctx, cudaSetDevice(ctx->device));
    FriLayerDev nxt;
    int rc = fri_fold_step(f, zeta, false, &nxt);
    if (rc != S252_OK) { dfree(ctx, nxt.evals); return rc; }
    // fri_last_value = coefficient 0 of the folded polynomial = mean of its evaluations on the remaining coset
    std::vector<fe> rem(nxt.size);
    cudaError_t e = cudaMemcpyAsync(rem.data(), nxt.evals, nxt.size * sizeof(fe), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    dfree(ctx, nxt.evals);
    CU(ctx, e);
    fe acc = H::zero();
    for (auto& v : rem) acc = H::add(acc, v);
    H::to_lw(H::mul(acc, H::inv(H::from_u64((uint64_t)nxt.size))), last_value->limbs);
    return S252_OK;
}
// fri_commit_phase continued from layer `layer_index` (a sharded commit phase collapses to one GPU once the
// layers are small): evals = the layer_size evaluations of that layer on the coset h^(2^layer_index) <w_layer_size>,
// resident on this device (internal format); number_layers = layers still to commit, this one included.  The
// transcript must be in the state the reference's is in before it appends this layer's root.
extern "C" int s252_fri_commit_phase_from_layer(s252_ctx* ctx, size_t number_layers, const void* evals, size_t layer_size,
                                                s252_transcript* transcript, uint64_t coset_offset, size_t layer_index, s252_fri** out,
                                                s252_fe* last_value, uint8_t* roots_out) {
    NVTX_RANGE("s252_fri_commit_phase_from_layer");
    if (!ctx || !evals || !transcript || !out || !last_value) return S252_ERR_INVALID;
    *out = nullptr;
    CU(ctx, cudaSetDevice(ctx->device));
    if (!is_pow2(layer_size) || number_layers > ilog2(layer_size) || coset_offset == 0 || layer_index >= 64) FAIL(ctx, S252_ERR_INVALID, "bad domain sizes for FRI");
    fe hk = H::from_u64(coset_offset);
    for (size_t j = 0; j < layer_index; ++j) hk = H::sqr(hk);
    s252_fri* f = new s252_fri();
    f->ctx = ctx; f->domain_size = layer_size;
    int rc = [&]() -> int {
        FriLayerDev cur;
        cur.size = layer_size;
        TRY(dalloc(ctx, &cur.evals, layer_size));
        f->layers.push_back(cur);
        CU(ctx, cudaMemcpyAsync(f->layers[0].evals, evals, layer_size * sizeof(fe), cudaMemcpyDeviceToDevice, ctx->stream));
        return fri_from_layer0(ctx, f, number_layers, transcript, hk, layer_size, last_value, roots_out);
    }();
    if (rc != S252_OK) { fri_free(f); return rc; }
    *out = f;
    return S252_OK;
}
// One fold of the commit phase on a block of rows (fri/mod.rs:43-51 as the verifier's formula, verifier.rs:511-512):
// out[j] = (v[j] + s[j])/2 + zeta (v[j] - s[j]) / (2 h_k w^(i0+j)),  v[j] = layer_k[i0 + j], s[j] = layer_k[i0 + j + layer_size/2],
// h_k = coset_offset^(2^k), w of order layer_size = domain_size >> k.  v, s, out: device, internal format, `count` elements.
extern "C" int s252_fri_fold_rows(s252_ctx* ctx, const void* v, const void* s, size_t count, size_t i0, size_t layer_size,
                                  size_t domain_size, size_t layer_index, const s252_fe* zeta, uint64_t coset_offset, void* out) {
    NVTX_RANGE("s252_fri_fold_rows");
    if (!ctx || !v || !s || !zeta || !out) return S252_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    if (!is_pow2(layer_size) || layer_size < 2 || i0 + count > layer_size / 2 || coset_offset == 0 || layer_index >= 64 ||
        (layer_size << layer_index) != domain_size)
        FAIL(ctx, S252_ERR_INVALID, "bad FRI fold block");
    if (count == 0) return S252_OK;
    fe hk = H::from_u64(coset_offset);
    for (size_t j = 0; j < layer_index; ++j) hk = H::sqr(hk);
    fe w;
    H::primitive_root(ilog2(domain_size), &w);
    const fe* inv_tw;                                       // w_domain^(-i): the table the single-GPU path uses, stride domain/layer
    TRY(get_power_table(ctx, domain_size / 2, H::inv(w), H::one(), &inv_tw));
    const fe inv2 = H::inv(H::from_u64(2));
    const fe cfac = H::mul(H::from_lw(zeta->limbs), H::mul(inv2, H::inv(hk)));
    prof_begin(ctx, "fri_fold_rows");
    prof_work(ctx, 96.0 * count, 3.2 * count, 0);
    s252::fri_fold_rows_kernel<<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(
        reinterpret_cast<const fe*>(v), reinterpret_cast<const fe*>(s), count, inv_tw, (unsigned long long)(domain_size / layer_size), i0, cfac,
        inv2, reinterpret_cast<fe*>(out));
    LAUNCH_CHECK(ctx);
    return S252_OK;
}
extern "C" void s252_fri_destroy(s252_fri* f) {
    if (!f) return;
    cudaSetDevice(f->ctx->device);
    fri_free(f);
}
extern "C" size_t s252_fri_n_layers(const s252_fri* f) { return f->layers.size(); }
extern "C" int s252_fri_read_layer(s252_fri* f, size_t layer, size_t first, size_t count, s252_fe* out) {
    s252_ctx* ctx = f->ctx;
    CU(ctx, cudaSetDevice(ctx->device));
    if (layer >= f->layers.size() || first + count > f->layers[layer].size) FAIL(ctx, S252_ERR_RANGE, "FRI layer read out of range");
    if (count == 0) return S252_OK;
    return read_internal_as_lw(ctx, f->layers[layer].evals + first, count, out);
}
extern "C" int s252_fri_read_nodes(s252_fri* f, size_t layer, size_t first, size_t count, uint8_t* out) {
    s252_ctx* ctx = f->ctx;
    CU(ctx, cudaSetDevice(ctx->device));
    if (layer >= f->layers.size() || first + count > 2 * f->layers[layer].size - 1) FAIL(ctx, S252_ERR_RANGE, "FRI node read out of range");
    CU(ctx, cudaMemcpyAsync(out, f->layers[layer].nodes + 4 * first, count * 32, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return S252_OK;
}
extern "C" int s252_fri_query(s252_fri* f, const uint64_t* iotas, size_t n_queries, s252_fe* evals, s252_fe* evals_sym,
                              uint8_t* paths, uint8_t* paths_sym, size_t path_stride) {
    NVTX_RANGE("s252_fri_query");
    s252_ctx* ctx = f->ctx;
    CU(ctx, cudaSetDevice(ctx->device));
    if (!iotas && n_queries) return S252_ERR_INVALID;
    const size_t L = f->layers.size(), Q = n_queries;
    if (L == 0 || Q == 0) return S252_OK;
    const unsigned logd = ilog2(f->domain_size);
    if ((paths || paths_sym) && path_stride < logd) FAIL(ctx, S252_ERR_INVALID, "path_stride %zu is smaller than the tree depth %u", path_stride, logd);
    const size_t stride = path_stride ? path_stride : logd;
    // one launch gathers every layer's values and authentication paths (both the index and its symmetric)
    std::vector<s252::FriLayerRef> refs(L);
    for (size_t k = 0; k < L; ++k) {
        if (f->layers[k].size != (f->domain_size >> k)) FAIL(ctx, S252_ERR_INVALID, "internal: unexpected FRI layer size");
        refs[k] = {f->layers[k].evals, f->layers[k].nodes};
    }
    Tmp<s252::FriLayerRef> drefs(ctx);
    Tmp<unsigned long long> didx(ctx);
    Tmp<fe> dvals(ctx);
    Tmp<uint64_t> dpaths(ctx);
    const size_t entries = 2 * Q * L;
    TRY(dalloc(ctx, &drefs.p, L));
    TRY(dalloc(ctx, &didx.p, Q));
    TRY(dalloc(ctx, &dvals.p, entries));
    TRY(dalloc(ctx, &dpaths.p, entries * stride * 4));
    CU(ctx, cudaMemcpyAsync(drefs.p, refs.data(), L * sizeof(s252::FriLayerRef), cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemcpyAsync(didx.p, iotas, Q * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemsetAsync(dpaths.p, 0, entries * stride * 32, ctx->stream));
    const unsigned long long total = (unsigned long long)entries * (stride + 1);
    prof_begin(ctx, "fri_query_gather");
    s252::fri_query_gather<<<(unsigned)((total + 127) / 128), 128, 0, ctx->stream>>>(drefs.p, (unsigned)L, logd, didx.p, (unsigned)Q,
                                                                                      (unsigned)stride, dvals.p, dpaths.p);
    LAUNCH_CHECK(ctx);
    std::vector<s252_fe> hv(entries);
    std::vector<uint8_t> hp(entries * stride * 32);
    CU(ctx, cudaMemcpyAsync(hv.data(), dvals.p, entries * sizeof(fe), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaMemcpyAsync(hp.data(), dpaths.p, hp.size(), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    const size_t half = Q * L;
    if (evals) std::memcpy(evals, hv.data(), half * sizeof(s252_fe));
    if (evals_sym) std::memcpy(evals_sym, hv.data() + half, half * sizeof(s252_fe));
    if (paths) std::memcpy(paths, hp.data(), half * stride * 32);
    if (paths_sym) std::memcpy(paths_sym, hp.data() + half * stride * 32, half * stride * 32);
    return S252_OK;
}

// --------------------------------------------------------------------------------------------
// grinding
// One round of the search: the window [base, base + 2^32) in batches of 2^18 nonces, of which this GPU takes
// batches part, part + parts, ...  *found = smallest accepted nonce of this GPU's batches (~0: none).
static const unsigned GRIND_GRID = 1024;                       // x 256 threads = 2^18 nonces per batch
static const unsigned GRIND_BATCHES = 1u << 14;                // per round, over all parts
static int grind_round(s252_ctx* ctx, const uint8_t challenge[32], uint8_t grinding_factor, uint64_t base, uint64_t limit,
                       unsigned part, unsigned parts, uint64_t* found, unsigned window_log = 32) {
    uint64_t lanes[4];
    std::memcpy(lanes, challenge, 32);   // little-endian host
    Tmp<unsigned long long> best(ctx);
    TRY(dalloc(ctx, &best.p, 1));
    CU(ctx, cudaMemsetAsync(best.p, 0xff, 8, ctx->stream));
    const uint64_t window = 1ull << window_log;
    const uint64_t window_end = base + window > base ? std::min<uint64_t>(limit, base + window) : limit;
    const unsigned batches = (unsigned)std::max<uint64_t>(1, window >> 18);          // 2^18 nonces per batch
    prof_begin(ctx, "grind_kernel");
    s252::grind_kernel<<<GRIND_GRID, 256, 0, ctx->stream>>>(lanes[0], lanes[1], lanes[2], lanes[3], base, window_end,
                                                            (batches + parts - 1) / parts, part, parts, grinding_factor, best.p);
    LAUNCH_CHECK(ctx);
    unsigned long long f;
    CU(ctx, cudaMemcpyAsync(&f, best.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    *found = f;
    return S252_OK;
}
extern "C" int s252_generate_nonce_with_grinding(s252_ctx* ctx, const uint8_t challenge[32], uint8_t grinding_factor,
                                                 uint64_t limit, uint64_t* nonce) {
    NVTX_RANGE("s252_generate_nonce_with_grinding");
    if (!ctx || !challenge || !nonce) return S252_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    if (grinding_factor > 64) FAIL(ctx, S252_ERR_NOT_FOUND, "a 64-bit head cannot have %u trailing zeros", grinding_factor);
    if (limit == 0) limit = ~0ull;   // the reference searches 0..u64::MAX
    // One persistent launch per 2^32-nonce window: the grid walks it in batches and drains as soon as every nonce
    // below the best hit has been tested (grind_kernel); factor 20 finishes in ~4 batches of the first window.
    uint64_t base = 0;
    while (base < limit) {
        uint64_t found;
        TRY(grind_round(ctx, challenge, grinding_factor, base, limit, 0, 1, &found));
        if (found != ~0ull) { *nonce = found; return S252_OK; }
        if (base + (1ull << 32) <= base) break;               // wrapped around 2^64
        base += 1ull << 32;
    }
    FAIL(ctx, S252_ERR_NOT_FOUND, "nonce not found below %llu", (unsigned long long)limit);
}
// The same search shared by `parts` GPUs (disjoint nonce ranges + a `min` all-reduce): one round over the
// window [base, base + 2^window_log) of which this GPU tests its round-robin share of the 2^18-nonce batches.  *found is
// this GPU's smallest accepted nonce or UINT64_MAX; the caller takes the minimum over the GPUs and, if none found
// anything, calls again with the next window.  A GPU does not see the others' hits while its kernel runs, so the window
// should be about twice the expected position of the first hit (2^(factor+1)): large windows make every GPU search until
// its OWN first hit.  The minimum over the parts is the reference's nonce (grinding.rs:44-47).
extern "C" int s252_grind_round(s252_ctx* ctx, const uint8_t challenge[32], uint8_t grinding_factor, uint64_t base, uint64_t limit,
                                unsigned part, unsigned parts, unsigned window_log, uint64_t* found) {
    NVTX_RANGE("s252_grind_round");
    if (!ctx || !challenge || !found || parts == 0 || part >= parts || window_log < 18 || window_log > 40) return S252_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    if (grinding_factor > 64) FAIL(ctx, S252_ERR_NOT_FOUND, "a 64-bit head cannot have %u trailing zeros", grinding_factor);
    if (limit == 0) limit = ~0ull;
    return grind_round(ctx, challenge, grinding_factor, base, limit, part, parts, found, window_log);
}

// --------------------------------------------------------------------------------------------
// host Keccak-256 (a few digests: the top levels above per-GPU subtree roots)
// ByteConversion::to_bytes_be for n elements (host): LW Montgomery -> 32-byte big-endian canonical values
extern "C" void s252_fe_to_bytes_be(const s252_fe* in, size_t n, uint8_t* out) {
    for (size_t i = 0; i < n; ++i) H::to_bytes_be(H::from_lw(in[i].limbs), out + 32 * i);
}
extern "C" void s252_keccak256(const uint8_t* data, size_t len, uint8_t out[32]) {
    H::Keccak256 k;
    k.update(data, len);
    k.finalize(out);
}

// --------------------------------------------------------------------------------------------
// transcript
extern "C" s252_transcript* s252_transcript_new(void) { return new s252_transcript(); }
extern "C" void s252_transcript_free(s252_transcript* t) { delete t; }
extern "C" void s252_transcript_append(s252_transcript* t, const uint8_t* data, size_t len) { t->append(data, len); }
extern "C" void s252_transcript_challenge(s252_transcript* t, uint8_t out[32]) { t->challenge(out); }
extern "C" void s252_transcript_to_field(s252_transcript* t, s252_fe* out) { H::to_lw(t->to_field(), out->limbs); }
extern "C" uint64_t s252_transcript_to_usize(s252_transcript* t) { return t->to_usize(); }

// --------------------------------------------------------------------------------------------
// diagnostics
extern "C" int s252_fe_binop(s252_ctx* ctx, int op, const s252_fe* a, const s252_fe* b, s252_fe* out, size_t n, int mem) {
    if (!ctx || !a || !out || op < 0 || op > 3 || (op != 3 && !b)) return S252_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    Tmp<fe> sa(ctx), sb(ctx), so(ctx);
    const fe *da, *db = nullptr;
    TRY(stage_in(ctx, a, n, mem, sa, &da));
    if (b) TRY(stage_in(ctx, b, n, mem, sb, &db));
    fe* dout = reinterpret_cast<fe*>(out);
    if (mem == S252_HOST) { TRY(dalloc(ctx, &so.p, n)); dout = so.p; }
    prof_begin(ctx, "fe_binop_kernel");
    s252::fe_binop_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(op, da, db ? db : da, dout, n);
    LAUNCH_CHECK(ctx);
    TRY(stage_out(ctx, dout, out, n, mem));
    return S252_OK;
}
extern "C" int s252_keccak256_batch(s252_ctx* ctx, const uint8_t* msgs, size_t msg_len, size_t n, uint8_t* digests) {
    if (!ctx || !digests || (!msgs && msg_len * n)) return S252_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    Tmp<uint8_t> dm(ctx), dd(ctx);
    TRY(dalloc(ctx, &dm.p, msg_len * n + 8));
    TRY(dalloc(ctx, &dd.p, 32 * n));
    if (msg_len * n) CU(ctx, cudaMemcpyAsync(dm.p, msgs, msg_len * n, cudaMemcpyHostToDevice, ctx->stream));
    prof_begin(ctx, "keccak_bytes_kernel");
    s252::keccak_bytes_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(dm.p, msg_len, n, dd.p);
    LAUNCH_CHECK(ctx);
    CU(ctx, cudaMemcpyAsync(digests, dd.p, 32 * n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return S252_OK;
}

template <typename F>
static int time_kernel(s252_ctx* ctx, F launch, int reps, float* ms_best) {
    cudaEvent_t e0, e1;
    CU(ctx, cudaEventCreate(&e0));
    CU(ctx, cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = 0; r < reps + 1; ++r) {
        CU(ctx, cudaEventRecord(e0, ctx->stream));
        launch();
        ctx->launches++;
        CU(ctx, cudaEventRecord(e1, ctx->stream));
        CU(ctx, cudaEventSynchronize(e1));
        float ms;
        CU(ctx, cudaEventElapsedTime(&ms, e0, e1));
        if (r > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    CU(ctx, cudaGetLastError());
    *ms_best = best;
    return S252_OK;
}
extern "C" int s252_microbench_int_pipes(s252_ctx* ctx, double out[8]) {
    if (!ctx || !out) return S252_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    int sms = 0;
    CU(ctx, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
    Tmp<uint32_t> sink(ctx);
    const int blocks = sms * 8, threads = 256, iters = 2048;
    TRY(dalloc(ctx, &sink.p, (size_t)blocks * threads));
    for (int which = 0; which < 8; ++which) {
        float ms;
        TRY(time_kernel(ctx, [&]() { s252::int_pipe_bench<<<blocks, threads, 0, ctx->stream>>>(which, iters, sink.p); }, 5, &ms));
        const double ops = (double)blocks * threads * iters * s252::INT_BENCH_OPS_PER_ITER;
        out[which] = ops / (ms * 1e-3) / 1e9;
    }
    return S252_OK;
}
extern "C" int s252_microbench_fe_mul(s252_ctx* ctx, double* gmuls) {
    if (!ctx || !gmuls) return S252_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    int sms = 0;
    CU(ctx, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
    Tmp<fe> sink(ctx);
    const int blocks = sms * 8, threads = 256, iters = 512;
    TRY(dalloc(ctx, &sink.p, (size_t)blocks * threads));
    float ms;
    TRY(time_kernel(ctx, [&]() { s252::fe_mul_bench<<<blocks, threads, 0, ctx->stream>>>(iters, sink.p); }, 5, &ms));
    *gmuls = (double)blocks * threads * iters * 2 / (ms * 1e-3) / 1e9;
    return S252_OK;
}
extern "C" int s252_microbench_keccak(s252_ctx* ctx, double* gperms) {
    if (!ctx || !gperms) return S252_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    int sms = 0;
    CU(ctx, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
    Tmp<uint64_t> sink(ctx);
    const int blocks = sms * 16, threads = 128, iters = 64;
    TRY(dalloc(ctx, &sink.p, (size_t)blocks * threads));
    float ms;
    TRY(time_kernel(ctx, [&]() { s252::keccak_bench<<<blocks, threads, 0, ctx->stream>>>(iters, sink.p); }, 5, &ms));
    *gperms = (double)blocks * threads * iters / (ms * 1e-3) / 1e9;
    return S252_OK;
}

// --------------------------------------------------------------------------------------------
// Cairo side (include/stark252_cairo.h)
#include "sharded.cuh"
#include "cairo_api.cuh"
#include "cairo_sharded.cuh"
