// sharded.cuh -- ONE interpolate_and_commit (src/starks/prover.rs:126-159) on the GPUs of one box, behind the C ABI: one process
// (or thread) per GPU, every rank calls the same entry point with its own columns, NCCL is called from here (SURVEY.md 8e rows
// 1, 3 and 7; the Python orchestration of distributed.py restated in C++ so that a non-Python caller can bind it).
//
//   columns   rank r owns the contiguous column range column_shards(c, G)[r]  (33 over 8 -> 5,4,4,..): iNTT + coset LDE of its own
//             columns, no communication (compute_lde_trace_evaluations, prover.rs:161-185);
//   exchange  "my columns, all rows" -> "all columns, my block of M/G rows": one ncclSend per (column, destination) straight out of
//             the LDE buffer, one ncclRecv per (column, source) straight into the receiver's column-major block; the rank's columns
//             are cut into pipeline groups and the exchange of group g runs on a second stream under the upload + LDE of group g+1;
//   tree      per-rank leaf hashing + subtree over its row block (BatchedMerkleTree::build, config.rs:19-20: the heap layout makes
//             an aligned block of M/G leaves a subtree), the G subtree roots all-gathered, the top log2(G) levels on every rank;
//   openings  the owner of a row serves its values and the subtree part of its path (one SUM all-reduce of a packed buffer: an
//             entry has one owner), the top levels are replicated (open_deep_composition_poly, prover.rs:484-529).
//
// NCCL is bound at run time (dlopen of libnccl.so.2, S252_NCCL_LIB overrides the name): the library has no link-time dependency on
// it and single-GPU users never load it.
#pragma once
#include <dlfcn.h>

namespace s252nccl {
typedef struct ncclComm* comm_t;
struct unique_id { char internal[128]; };
enum { kUint8 = 1, kUint64 = 5, kSum = 0, kMax = 2, kMin = 3 };
struct Api {
    void* lib = nullptr;
    int (*GetUniqueId)(unique_id*) = nullptr;
    int (*CommInitRank)(comm_t*, int, unique_id, int) = nullptr;
    int (*CommDestroy)(comm_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, comm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string err;
};
static Api& api() {
    static Api a;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* name = std::getenv("S252_NCCL_LIB");
        a.lib = dlopen(name && *name ? name : "libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!a.lib) { a.err = std::string("NCCL not found: ") + dlerror(); return; }
        auto sym = [&](const char* s) -> void* { void* p = dlsym(a.lib, s); if (!p && a.err.empty()) a.err = std::string("NCCL symbol missing: ") + s; return p; };
        a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
        a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
        a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
        a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(sym("ncclGroupStart"));
        a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(sym("ncclGroupEnd"));
        a.Send = reinterpret_cast<decltype(a.Send)>(sym("ncclSend"));
        a.Recv = reinterpret_cast<decltype(a.Recv)>(sym("ncclRecv"));
        a.AllGather = reinterpret_cast<decltype(a.AllGather)>(sym("ncclAllGather"));
        a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(sym("ncclAllReduce"));
        a.Broadcast = reinterpret_cast<decltype(a.Broadcast)>(sym("ncclBroadcast"));
        a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
    });
    return a;
}
}  // namespace s252nccl

struct s252_comm {
    s252_ctx* ctx = nullptr;
    s252nccl::comm_t comm = nullptr;
    int rank = 0, world = 1;
    cudaStream_t xstream = nullptr;        // the exchange runs here, under the next group's transforms on ctx->stream
    cudaEvent_t ready = nullptr, done = nullptr;
};

struct s252_sharded_commit {
    s252_comm* comm = nullptr;
    size_t n_rows = 0, n_cols = 0, rows_per = 0;      // M, all columns, M / G
    std::vector<s252_commit*> local;                  // this rank's columns (coefficients + LDE, all rows), one handle per pipeline group
    std::vector<size_t> local_cols;                   // columns per group
    size_t col0 = 0;                                  // first column of this rank's shard
    fe* block_cols = nullptr;                         // [n_cols][rows_per]: all columns, this rank's rows
    s252_commit* block = nullptr;                     // the subtree over them (in place)
    std::vector<std::array<uint8_t, 32>> top;         // replicated heap over the G subtree roots, root at 0
};

#define NCCL_TRY(ctx, call)                                                                                              \
    do {                                                                                                                 \
        int _r = (call);                                                                                                 \
        if (_r != 0) FAIL(ctx, S252_ERR_CUDA, "%s: %s", #call, s252nccl::api().GetErrorString ? s252nccl::api().GetErrorString(_r) : "?"); \
    } while (0)

// contiguous balanced ranges (the same rule on every rank: distributed.py column_shards)
static void shard_range(size_t n, size_t parts, size_t r, size_t* lo, size_t* hi) {
    const size_t base = n / parts, extra = n % parts;
    *lo = r * base + std::min(r, extra);
    *hi = *lo + base + (r < extra ? 1 : 0);
}

extern "C" int s252_comm_unique_id(uint8_t id[S252_COMM_ID_BYTES]) {
    if (!id) return S252_ERR_INVALID;
    auto& A = s252nccl::api();
    if (!A.err.empty() || !A.GetUniqueId) return S252_ERR_CUDA;
    s252nccl::unique_id u;
    if (A.GetUniqueId(&u) != 0) return S252_ERR_CUDA;
    std::memcpy(id, u.internal, S252_COMM_ID_BYTES);
    return S252_OK;
}
extern "C" int s252_comm_create(s252_ctx* ctx, const uint8_t id[S252_COMM_ID_BYTES], int rank, int world, s252_comm** out) {
    if (!ctx || !id || !out || world < 1 || rank < 0 || rank >= world) return S252_ERR_INVALID;
    *out = nullptr;
    if (world & (world - 1)) FAIL(ctx, S252_ERR_INVALID, "the number of ranks must be a power of two (row blocks must be subtrees), got %d", world);
    auto& A = s252nccl::api();
    if (!A.err.empty()) FAIL(ctx, S252_ERR_CUDA, "%s", A.err.c_str());
    CU(ctx, cudaSetDevice(ctx->device));
    s252_comm* c = new s252_comm();
    c->ctx = ctx; c->rank = rank; c->world = world;
    s252nccl::unique_id u;
    std::memcpy(u.internal, id, S252_COMM_ID_BYTES);
    const int r = A.CommInitRank(&c->comm, world, u, rank);
    if (r != 0) { delete c; FAIL(ctx, S252_ERR_CUDA, "ncclCommInitRank: %s", A.GetErrorString(r)); }
    // highest priority: the send/receive kernels must get SM slots while the transforms of the next group fill the GPU
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (cudaStreamCreateWithPriority(&c->xstream, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ready, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->done, cudaEventDisableTiming) != cudaSuccess) {
        A.CommDestroy(c->comm);
        delete c;
        FAIL(ctx, S252_ERR_CUDA, "stream/event creation failed");
    }
    *out = c;
    return S252_OK;
}
extern "C" void s252_comm_destroy(s252_comm* c) {
    if (!c) return;
    cudaSetDevice(c->ctx->device);
    cudaStreamSynchronize(c->xstream);
    cudaStreamSynchronize(c->ctx->stream);
    if (c->comm) s252nccl::api().CommDestroy(c->comm);
    cudaEventDestroy(c->ready);
    cudaEventDestroy(c->done);
    cudaStreamDestroy(c->xstream);
    delete c;
}
extern "C" int s252_comm_rank(const s252_comm* c) { return c ? c->rank : -1; }
extern "C" int s252_comm_world(const s252_comm* c) { return c ? c->world : 0; }

extern "C" void s252_sharded_commit_destroy(s252_sharded_commit* sc) {
    if (!sc) return;
    s252_ctx* ctx = sc->comm->ctx;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(sc->comm->xstream);          // nothing of an exchange may still read the LDE buffers
    if (sc->block) commit_free(sc->block);
    dfree(ctx, sc->block_cols);
    for (s252_commit* h : sc->local) commit_free(h);
    delete sc;
}

// rows [0, width) of `height` columns from a table with column pitch spitch into one with pitch dpitch (elements); 16 bytes per
// thread, grid-stride.  The DMA engine's 2-D device-to-device copy runs at a fraction of HBM speed; at one or two ranks the rows a
// rank keeps for itself are most of the table.
__global__ void __launch_bounds__(256) copy_rows_kernel(uint4* __restrict__ dst, unsigned long long dpitch, const uint4* __restrict__ src,
                                                        unsigned long long spitch, unsigned long long width, unsigned long long height) {
    const unsigned long long per_col = 2 * width, total = per_col * height;       // an element is two 16-byte halves
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long c = i / per_col, r = i - c * per_col;
        dst[c * 2 * dpitch + r] = src[c * 2 * spitch + r];
    }
}
static int copy_rows(s252_ctx* ctx, fe* dst, size_t dpitch, const fe* src, size_t spitch, size_t width, size_t height) {
    if (width == 0 || height == 0) return S252_OK;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
    const unsigned long long total = 2ull * width * height;
    const unsigned blocks = (unsigned)std::min<unsigned long long>((total + 255) / 256, (unsigned long long)sms * 16);
    prof_begin(ctx, "copy_rows_kernel");
    prof_work(ctx, 64.0 * width * height, 0, 0);
    copy_rows_kernel<<<blocks, 256, 0, ctx->stream>>>(reinterpret_cast<uint4*>(dst), dpitch, reinterpret_cast<const uint4*>(src), spitch, width, height);
    LAUNCH_CHECK(ctx);
    return S252_OK;
}

// The tail of every row-block commit: leaves + subtree over sc->block_cols in place, subtree roots all-gathered device to device,
// top levels on the host of every rank.
static int sharded_finish_tree(s252_ctx* ctx, s252_comm* comm, s252_sharded_commit* sc, uint8_t root[32]) {
    auto& A = s252nccl::api();
    const size_t G = (size_t)comm->world;
    TRY(s252_commit_device_columns_inplace(ctx, sc->block_cols, sc->rows_per, sc->n_cols, sc->rows_per, &sc->block, nullptr));
    Tmp<uint64_t> droots(ctx);
    TRY(dalloc(ctx, &droots.p, 4 * G));
    if (G > 1) NCCL_TRY(ctx, A.AllGather(sc->block->nodes, droots.p, 32, s252nccl::kUint8, comm->comm, ctx->stream));
    else CU(ctx, cudaMemcpyAsync(droots.p, sc->block->nodes, 32, cudaMemcpyDeviceToDevice, ctx->stream));
    std::vector<uint8_t> roots(32 * G);
    CU(ctx, cudaMemcpyAsync(roots.data(), droots.p, 32 * G, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    // heap over the G subtree roots, node = Keccak256(left || right)
    sc->top.assign(2 * G - 1, std::array<uint8_t, 32>{});
    for (size_t g = 0; g < G; ++g) std::memcpy(sc->top[G - 1 + g].data(), roots.data() + 32 * g, 32);
    for (size_t i = G - 1; i-- > 0;) {
        H::Keccak256 k;
        k.update(sc->top[2 * i + 1].data(), 32);
        k.update(sc->top[2 * i + 2].data(), 32);
        k.finalize(sc->top[i].data());
    }
    std::memcpy(root, sc->top[0].data(), 32);
    return S252_OK;
}

// batch_commit (prover.rs:96-104) of a table whose ROWS are already spread over the ranks in equal contiguous blocks (the (H1, H2)
// table, a FRI layer): block_cols [n_cols][M/G], device, internal format; the handle takes ownership of the buffer.
static int sharded_commit_row_block(s252_ctx* ctx, s252_comm* comm, fe* block_cols, size_t n_cols, size_t M, s252_sharded_commit** out,
                                    uint8_t root[32]) {
    s252_sharded_commit* sc = new s252_sharded_commit();
    sc->comm = comm; sc->n_rows = M; sc->n_cols = n_cols; sc->rows_per = M / (size_t)comm->world;
    sc->block_cols = block_cols;
    const int rc = sharded_finish_tree(ctx, comm, sc, root);
    if (rc != S252_OK) { s252_sharded_commit_destroy(sc); return rc; }
    *out = sc;
    return S252_OK;
}

// The column-sharded commit around a producer of LDE handles: producer(g, &h) queues the upload + iNTT + coset LDE of this rank's
// pipeline group g (group_cols[g] columns, M rows) on ctx->stream and returns its handle.
template <typename Producer>
static int sharded_commit_core(s252_ctx* ctx, s252_comm* comm, const size_t* group_cols, size_t n_groups, size_t M, size_t n_cols_total,
                               Producer&& producer, s252_sharded_commit** out, uint8_t root[32]) {
    *out = nullptr;
    auto& A = s252nccl::api();
    const size_t G = (size_t)comm->world, me = (size_t)comm->rank;
    if (n_groups == 0 || n_groups > S252_MAX_PIPELINE_GROUPS) FAIL(ctx, S252_ERR_INVALID, "between 1 and %d pipeline groups", S252_MAX_PIPELINE_GROUPS);
    if (M % G || M / G < 1) FAIL(ctx, S252_ERR_INVALID, "more ranks than LDE rows");
    const size_t rows_per = M / G;
    size_t lo, hi, mine = 0;
    shard_range(n_cols_total, G, me, &lo, &hi);
    for (size_t g = 0; g < n_groups; ++g) { if (group_cols[g] == 0) return S252_ERR_INVALID; mine += group_cols[g]; }
    if (mine != hi - lo) FAIL(ctx, S252_ERR_INVALID, "rank %zu of %zu holds columns [%zu, %zu) of %zu, got %zu", me, G, lo, hi, n_cols_total, mine);

    s252_sharded_commit* sc = new s252_sharded_commit();
    sc->comm = comm; sc->n_rows = M; sc->n_cols = n_cols_total; sc->rows_per = rows_per; sc->col0 = lo;
    int rc = [&]() -> int {
        // every rank's group widths (a rank may have fewer groups than another: zero-width groups send nothing)
        std::vector<uint64_t> counts(G * S252_MAX_PIPELINE_GROUPS, 0), my(S252_MAX_PIPELINE_GROUPS, 0);
        for (size_t g = 0; g < n_groups; ++g) my[g] = group_cols[g];
        if (G > 1) {
            Tmp<uint64_t> dcounts(ctx);
            TRY(dalloc(ctx, &dcounts.p, (G + 1) * S252_MAX_PIPELINE_GROUPS));
            uint64_t* dmy = dcounts.p + G * S252_MAX_PIPELINE_GROUPS;
            CU(ctx, cudaMemcpyAsync(dmy, my.data(), my.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
            NCCL_TRY(ctx, A.AllGather(dmy, dcounts.p, S252_MAX_PIPELINE_GROUPS, s252nccl::kUint64, comm->comm, ctx->stream));
            CU(ctx, cudaMemcpyAsync(counts.data(), dcounts.p, counts.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
            CU(ctx, cudaStreamSynchronize(ctx->stream));
        } else {
            counts = my;
        }
        size_t rounds = 0;
        for (size_t r = 0; r < G; ++r) {
            size_t rlo, rhi, tot = 0, ng = 0;
            shard_range(n_cols_total, G, r, &rlo, &rhi);
            for (size_t g = 0; g < S252_MAX_PIPELINE_GROUPS; ++g) { tot += counts[r * S252_MAX_PIPELINE_GROUPS + g]; if (counts[r * S252_MAX_PIPELINE_GROUPS + g]) ng = g + 1; }
            if (tot != rhi - rlo) FAIL(ctx, S252_ERR_INVALID, "rank %zu announced %zu columns, its shard has %zu", r, tot, rhi - rlo);
            rounds = std::max(rounds, ng);
        }
        TRY(dalloc(ctx, &sc->block_cols, n_cols_total * rows_per));
        std::vector<size_t> sent(G, 0);                            // columns of rank r's shard already exchanged
        for (size_t g = 0; g < rounds; ++g) {
            const size_t cg = g < n_groups ? group_cols[g] : 0;
            s252_commit* h = nullptr;
            if (cg) {
                TRY(producer(g, &h));
                sc->local.push_back(h);
                sc->local_cols.push_back(cg);
                // my own rows of these columns stay on this GPU
                TRY(copy_rows(ctx, sc->block_cols + (lo + sent[me]) * rows_per, rows_per, h->lde + me * rows_per, M, rows_per, cg));
            }
            if (G > 1) {
                // the exchange of this group starts when its LDE is complete and runs beside the next group's upload + transforms
                CU(ctx, cudaEventRecord(comm->ready, ctx->stream));
                CU(ctx, cudaStreamWaitEvent(comm->xstream, comm->ready, 0));
                NCCL_TRY(ctx, A.GroupStart());
                for (size_t d = 0; d < G; ++d) {
                    if (d == me) continue;
                    for (size_t c = 0; c < cg; ++c)
                        NCCL_TRY(ctx, A.Send(h->lde + c * M + d * rows_per, rows_per * sizeof(fe), s252nccl::kUint8, (int)d, comm->comm, comm->xstream));
                }
                for (size_t s = 0; s < G; ++s) {
                    if (s == me) continue;
                    size_t slo, shi;
                    shard_range(n_cols_total, G, s, &slo, &shi);
                    const size_t sc_g = counts[s * S252_MAX_PIPELINE_GROUPS + g];
                    for (size_t c = 0; c < sc_g; ++c)
                        NCCL_TRY(ctx, A.Recv(sc->block_cols + (slo + sent[s] + c) * rows_per, rows_per * sizeof(fe), s252nccl::kUint8, (int)s, comm->comm, comm->xstream));
                }
                NCCL_TRY(ctx, A.GroupEnd());
            }
            for (size_t r = 0; r < G; ++r) sent[r] += counts[r * S252_MAX_PIPELINE_GROUPS + g];
        }
        if (G > 1) {
            CU(ctx, cudaEventRecord(comm->done, comm->xstream));
            CU(ctx, cudaStreamWaitEvent(ctx->stream, comm->done, 0));
        }
        return sharded_finish_tree(ctx, comm, sc, root);
    }();
    if (rc != S252_OK) {
        cudaStreamSynchronize(comm->xstream);
        s252_sharded_commit_destroy(sc);
        return rc;
    }
    *out = sc;
    return S252_OK;
}

// group_tables[g]: row-major TraceTable [n_rows][group_cols[g]] of this rank's pipeline group g (host or device, `mem`); the groups
// in order are this rank's column shard.  Collective: every rank of the communicator calls it with the same shape arguments.
extern "C" int s252_interpolate_and_commit_sharded(s252_ctx* ctx, s252_comm* comm, const s252_fe* const* group_tables, const size_t* group_cols,
                                                   size_t n_groups, size_t n_rows, size_t n_cols_total, size_t blowup, uint64_t coset_offset,
                                                   int mem, s252_sharded_commit** out, uint8_t root[32]) {
    NVTX_RANGE("s252_interpolate_and_commit_sharded");
    if (!ctx || !comm || comm->ctx != ctx || !group_tables || !group_cols || !out || !root || n_groups == 0 || n_groups > S252_MAX_PIPELINE_GROUPS)
        return S252_ERR_INVALID;
    *out = nullptr;
    CU(ctx, cudaSetDevice(ctx->device));
    if (!is_pow2(n_rows) || !is_pow2(blowup) || blowup > MAX_COSETS || coset_offset == 0) FAIL(ctx, S252_ERR_INVALID, "bad LDE shape");
    for (size_t g = 0; g < n_groups; ++g) if (!group_tables[g]) return S252_ERR_INVALID;
    // host tables: group g+1 is uploaded on the copy stream (one contiguous DMA into a staging buffer) while group g is
    // interpolated and extended; the staging buffers come from the arena, so the copy stream first waits for whatever the
    // compute stream still does with recycled blocks
    struct StagingSet {
        s252_ctx* ctx;
        std::vector<fe*> p;
        ~StagingSet() { for (fe* q : p) dfree(ctx, q); }
    } staging{ctx, {}};
    auto upload = [&](size_t g) -> int {
        CU(ctx, cudaMemcpyAsync(staging.p[g], group_tables[g], n_rows * group_cols[g] * sizeof(fe), cudaMemcpyHostToDevice, ctx->copy_stream));
        return S252_OK;
    };
    if (mem == S252_HOST) {
        for (size_t g = 0; g < n_groups; ++g) { fe* q = nullptr; TRY(dalloc(ctx, &q, n_rows * group_cols[g])); staging.p.push_back(q); }
        CU(ctx, cudaEventRecord(comm->ready, ctx->stream));
        CU(ctx, cudaStreamWaitEvent(ctx->copy_stream, comm->ready, 0));
        TRY(upload(0));
    }
    auto producer = [&](size_t g, s252_commit** h) -> int {
        const s252_fe* table = group_tables[g];
        if (mem == S252_HOST) {
            TRY(s252_copy_stream_wait(ctx));               // the transforms of group g wait for its upload ..
            if (g + 1 < n_groups) TRY(upload(g + 1));      // .. and run beside the next one
            table = reinterpret_cast<const s252_fe*>(staging.p[g]);
        }
        return interpolate_lde_impl(ctx, table, n_rows, group_cols[g], blowup, coset_offset, S252_DEVICE, false, h, nullptr);
    };
    return sharded_commit_core(ctx, comm, group_cols, n_groups, n_rows * blowup, n_cols_total, producer, out, root);
}

extern "C" size_t s252_sharded_commit_n_rows(const s252_sharded_commit* sc) { return sc->n_rows; }
extern "C" size_t s252_sharded_commit_n_cols(const s252_sharded_commit* sc) { return sc->n_cols; }
extern "C" size_t s252_sharded_commit_n_local(const s252_sharded_commit* sc) { return sc->local.size(); }
extern "C" s252_commit* s252_sharded_commit_local(const s252_sharded_commit* sc, size_t group) { return group < sc->local.size() ? sc->local[group] : nullptr; }
extern "C" s252_commit* s252_sharded_commit_block(const s252_sharded_commit* sc) { return sc->block; }

// The openings of several row-block commits of one communicator with ONE reduction (distributed.py open_many_packed): every rank
// writes the rows and subtree paths of the positions it owns into a zero-filled buffer with a fixed layout (an entry has exactly one
// owner), the buffers are summed, and the replicated top levels are appended.  rows_out[i]: [n_i][cols_i] LW; paths_out[i]:
// [n_i][log2(rows_i)][32] leaf -> root.  The result is complete on every rank.
static int sharded_open_many(s252_ctx* ctx, s252_comm* comm, s252_sharded_commit* const* commits, const std::vector<uint64_t>* index_lists,
                             size_t n_commits, std::vector<s252_fe>* rows_out, std::vector<uint8_t>* paths_out) {
    auto& A = s252nccl::api();
    const size_t G = (size_t)comm->world, me = (size_t)comm->rank;
    const unsigned depth_top = ilog2(G);
    std::vector<size_t> off(n_commits + 1, 0), rec(n_commits);
    for (size_t i = 0; i < n_commits; ++i) {
        const s252_sharded_commit* sc = commits[i];
        for (uint64_t q : index_lists[i])
            if (q >= sc->n_rows) FAIL(ctx, S252_ERR_RANGE, "position %llu out of range (%zu rows)", (unsigned long long)q, sc->n_rows);
        rec[i] = 32 * (sc->n_cols + ilog2(sc->rows_per));
        off[i + 1] = off[i] + rec[i] * index_lists[i].size();
    }
    std::vector<uint8_t> host(off[n_commits], 0);
    for (size_t i = 0; i < n_commits; ++i) {
        s252_sharded_commit* sc = commits[i];
        const size_t c = sc->n_cols;
        const unsigned depth_sub = ilog2(sc->rows_per);
        std::vector<uint64_t> mine_idx;
        std::vector<size_t> mine_q;
        for (size_t q = 0; q < index_lists[i].size(); ++q)
            if (index_lists[i][q] / sc->rows_per == me) { mine_q.push_back(q); mine_idx.push_back(index_lists[i][q] % sc->rows_per); }
        if (mine_q.empty()) continue;
        std::vector<s252_fe> rows(mine_q.size() * c);
        std::vector<uint8_t> paths(mine_q.size() * std::max(depth_sub, 1u) * 32);
        TRY(s252_commit_open(sc->block, mine_idx.data(), mine_idx.size(), rows.data(), paths.data()));
        for (size_t k = 0; k < mine_q.size(); ++k) {
            uint8_t* r = host.data() + off[i] + rec[i] * mine_q[k];
            std::memcpy(r, rows.data() + k * c, 32 * c);
            std::memcpy(r + 32 * c, paths.data() + k * std::max(depth_sub, 1u) * 32, 32 * depth_sub);
        }
    }
    if (G > 1 && !host.empty()) {
        Tmp<uint8_t> d(ctx);
        TRY(dalloc(ctx, &d.p, host.size()));
        CU(ctx, cudaMemcpyAsync(d.p, host.data(), host.size(), cudaMemcpyHostToDevice, ctx->stream));
        NCCL_TRY(ctx, A.AllReduce(d.p, d.p, host.size(), s252nccl::kUint8, s252nccl::kSum, comm->comm, ctx->stream));
        CU(ctx, cudaMemcpyAsync(host.data(), d.p, host.size(), cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
    }
    for (size_t i = 0; i < n_commits; ++i) {
        const s252_sharded_commit* sc = commits[i];
        const size_t c = sc->n_cols, n = index_lists[i].size();
        const unsigned depth_sub = ilog2(sc->rows_per);
        const size_t depth = depth_sub + depth_top;
        rows_out[i].assign(n * c, s252_fe{});
        paths_out[i].assign(n * std::max<size_t>(depth, 1) * 32, 0);
        for (size_t q = 0; q < n; ++q) {
            const uint8_t* r = host.data() + off[i] + rec[i] * q;
            std::memcpy(rows_out[i].data() + q * c, r, 32 * c);
            uint8_t* p = paths_out[i].data() + q * depth * 32;
            std::memcpy(p, r + 32 * c, 32 * depth_sub);
            size_t node = (G - 1) + index_lists[i][q] / sc->rows_per;     // heap index of the owner's subtree root
            for (unsigned l = 0; node != 0; ++l) {
                const size_t sib = (node & 1) ? node + 1 : node - 1;
                std::memcpy(p + 32 * (depth_sub + l), sc->top[sib].data(), 32);
                node = (node - 1) >> 1;
            }
        }
    }
    return S252_OK;
}

// Rows and authentication paths (leaf -> root, log2(n_rows) digests) of global positions, on EVERY rank.  Collective.
extern "C" int s252_sharded_commit_open(s252_sharded_commit* sc, const uint64_t* indices, size_t n_idx, s252_fe* rows_out, uint8_t* paths_out) {
    NVTX_RANGE("s252_sharded_commit_open");
    if (!sc || (!indices && n_idx) || !rows_out || !paths_out) return S252_ERR_INVALID;
    s252_ctx* ctx = sc->comm->ctx;
    CU(ctx, cudaSetDevice(ctx->device));
    if (n_idx == 0) return S252_OK;
    std::vector<uint64_t> idx(indices, indices + n_idx);
    std::vector<s252_fe> rows;
    std::vector<uint8_t> paths;
    s252_sharded_commit* one[1] = {sc};
    TRY(sharded_open_many(ctx, sc->comm, one, &idx, 1, &rows, &paths));
    std::memcpy(rows_out, rows.data(), rows.size() * sizeof(s252_fe));
    std::memcpy(paths_out, paths.data(), (size_t)ilog2(sc->n_rows) * 32 * n_idx);
    return S252_OK;
}
