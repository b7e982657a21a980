// cairo_sharded.cuh -- generate_cairo_proof (src/cairo/air.rs:1183-1190 = prove::<Stark252PrimeField, CairoAIR>, src/starks/prover.rs:
// 531-560) as ONE collective call over the GPUs of one box: every rank (one process or thread per GPU, each with its own context and
// its end of an s252_comm) calls s252_cairo_prove_sharded with the same trace and options; rank 0 gets StarkProof::serialize bytes,
// byte-identical to the single-GPU proof.  The partition is the one of lambdaworks_cairo_prover_b200/cairo_distributed.py (which
// remains the gloo-testable statement of the index logic); NCCL is called from here.
//
//   round 1  main and auxiliary trace: columns sharded for upload + iNTT + coset LDE, per-column ncclSend/ncclRecv to row blocks under
//            the next column group's transforms, per-rank subtree, roots all-gathered (sharded_commit_core).  The 11 main columns
//            build_auxiliary_trace reads are broadcast over NVLink by the ranks that hold them; every rank builds the auxiliary
//            columns (a global sort + scans) and keeps its own.
//   round 2  constraint evaluation on the rank's block of LDE rows (+ a `blowup`-row halo from the next rank), evaluations all-gathered,
//            H interpolated and (H1, H2) extended on every rank (2 columns), row-block tree.
//   round 3  out-of-domain evaluations of the rank's own columns, summed into place (an entry has one owner).
//   round 4  DEEP composition polynomial on the rank's rows; FRI by row blocks: per-layer row-block trees, pairwise half-layer exchange
//            per fold, collapse to rank 0 once a layer has at most 2^19 evaluations; grinding split over the ranks (MIN all-reduce);
//            openings served by the row owners in one reduction.
#pragma once

namespace {
struct ShardedFri {
    std::vector<s252_sharded_commit*> layers;      // layers 0 .. tail_first-1, row blocks
    s252_fri* tail = nullptr;                      // rank 0: layers tail_first .. L-1
    size_t tail_first = 0;
    std::vector<uint8_t> roots;                    // 32 bytes per layer, all layers
    s252_fe last{};
};
}  // namespace

// A FRI layer of at most 2^19 evaluations costs one GPU ~0.2 ms (fold + leaves + tree): sharding it saves less than the gather of the
// subtree roots and the fold exchange cost in latency.  S252_FRI_COLLAPSE_LOG overrides the threshold (the tests keep small layers
// sharded with it).
static unsigned sharded_fri_collapse_log() {
    if (const char* e = std::getenv("S252_FRI_COLLAPSE_LOG")) { const int v = std::atoi(e); if (v >= 1 && v <= 40) return (unsigned)v; }
    return 19;
}

// all ranks learn whether a rank-local step failed anywhere (instead of waiting forever in the next collective)
static int sharded_all_ok(s252_ctx* ctx, s252_comm* comm, int rc) {
    if (comm->world == 1) return rc;
    auto& A = s252nccl::api();
    Tmp<uint64_t> d(ctx);
    if (dalloc(ctx, &d.p, 1) != S252_OK) return S252_ERR_CUDA;
    const uint64_t mine = rc == S252_OK ? 0 : 1;
    uint64_t any = 1;
    if (cudaMemcpyAsync(d.p, &mine, 8, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
        A.AllReduce(d.p, d.p, 1, s252nccl::kUint64, s252nccl::kMax, comm->comm, ctx->stream) != 0 ||
        cudaMemcpyAsync(&any, d.p, 8, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess || cudaStreamSynchronize(ctx->stream) != cudaSuccess)
        return rc != S252_OK ? rc : S252_ERR_CUDA;
    if (rc != S252_OK) return rc;
    if (any) FAIL(ctx, S252_ERR_INVALID, "a peer rank failed in a rank-local step of the sharded prover");
    return S252_OK;
}

// fri_commit_phase (fri/mod.rs:20-72) with layer 0 given as this rank's block of the evaluations on the LDE coset; takes ownership of
// p0_block.  Every rank's transcript ends in the same state.
static int sharded_fri_commit_phase(s252_ctx* ctx, s252_comm* comm, fe* p0_block, size_t M, size_t layers, s252_transcript* t,
                                    uint64_t coset_offset, unsigned collapse_log, ShardedFri* F) {
    auto& A = s252nccl::api();
    const size_t G = (size_t)comm->world, me = (size_t)comm->rank;
    fe* block = p0_block;                          // owned until a layer commit adopts it
    size_t size = M, k = 0;
    F->roots.assign(32 * std::max<size_t>(layers, 1), 0);
    int rc = [&]() -> int {
        while (G > 1 && k + 1 < layers && size > ((size_t)1 << collapse_log) && size / G >= 2) {
            const size_t B = size / G, half = B / 2;
            s252_sharded_commit* sc = nullptr;
            uint8_t root[32];
            fe* cur = block;
            block = nullptr;                                                       // the handle owns `cur` from here on (also on failure)
            TRY(sharded_commit_row_block(ctx, comm, cur, 1, size, &sc, root));
            F->layers.push_back(sc);
            std::memcpy(F->roots.data() + 32 * k, root, 32);
            t->append(root, 32);
            s252_fe zeta;
            H::to_lw(t->to_field(), zeta.limbs);                                   // fri/mod.rs:41
            // the pairwise exchange before the fold: the partner of row i is i + size/2, G/2 ranks away
            Tmp<fe> v(ctx), s(ctx);
            TRY(dalloc(ctx, &v.p, half));
            TRY(dalloc(ctx, &s.p, half));
            const size_t base = 2 * (me % (G / 2));
            const bool am_v = me < G / 2;
            const size_t src_v = me / 2, src_s = me / 2 + G / 2;
            NCCL_TRY(ctx, A.GroupStart());
            for (int piece = 0; piece < 2; ++piece) {
                const size_t dst = base + piece;
                const fe* src = cur + piece * half;
                if (dst == me) CU(ctx, cudaMemcpyAsync(am_v ? v.p : s.p, src, half * sizeof(fe), cudaMemcpyDeviceToDevice, ctx->stream));
                else NCCL_TRY(ctx, A.Send(src, half * sizeof(fe), s252nccl::kUint8, (int)dst, comm->comm, ctx->stream));
            }
            if (src_v != me) NCCL_TRY(ctx, A.Recv(v.p, half * sizeof(fe), s252nccl::kUint8, (int)src_v, comm->comm, ctx->stream));
            if (src_s != me) NCCL_TRY(ctx, A.Recv(s.p, half * sizeof(fe), s252nccl::kUint8, (int)src_s, comm->comm, ctx->stream));
            NCCL_TRY(ctx, A.GroupEnd());
            fe* nxt = nullptr;
            TRY(dalloc(ctx, &nxt, half));
            block = nxt;
            TRY(s252_fri_fold_rows(ctx, v.p, s.p, half, me * half, size, M, k, &zeta, coset_offset, nxt));
            size /= 2;
            ++k;
        }
        // collapse: layer k in full, the rest of the phase on rank 0 (device-side transcript chain + tail kernel)
        Tmp<fe> full(ctx);
        const fe* layer = block;
        if (G > 1) {
            TRY(dalloc(ctx, &full.p, size));
            NCCL_TRY(ctx, A.AllGather(block, full.p, (size / G) * sizeof(fe), s252nccl::kUint8, comm->comm, ctx->stream));
            layer = full.p;
        }
        const size_t n_tail = layers - k;
        F->tail_first = k;
        std::vector<uint8_t> payload(32 * (n_tail + 1), 0);                        // tail roots + last value (LW bytes)
        int rc0 = S252_OK;
        if (me == 0) {
            s252_fe last;
            std::vector<uint8_t> tr(32 * std::max<size_t>(n_tail, 1));
            rc0 = s252_fri_commit_phase_from_layer(ctx, n_tail, layer, size, t, coset_offset, k, &F->tail, &last, tr.data());
            if (rc0 == S252_OK) { std::memcpy(payload.data(), tr.data(), 32 * n_tail); std::memcpy(payload.data() + 32 * n_tail, &last, 32); }
        }
        TRY(sharded_all_ok(ctx, comm, rc0));
        if (G > 1) {
            Tmp<uint8_t> d(ctx);
            TRY(dalloc(ctx, &d.p, payload.size()));
            CU(ctx, cudaMemcpyAsync(d.p, payload.data(), payload.size(), cudaMemcpyHostToDevice, ctx->stream));
            NCCL_TRY(ctx, A.Broadcast(d.p, d.p, payload.size(), s252nccl::kUint8, 0, comm->comm, ctx->stream));
            CU(ctx, cudaMemcpyAsync(payload.data(), d.p, payload.size(), cudaMemcpyDeviceToHost, ctx->stream));
            CU(ctx, cudaStreamSynchronize(ctx->stream));
        }
        std::memcpy(&F->last, payload.data() + 32 * n_tail, 32);
        std::memcpy(F->roots.data() + 32 * k, payload.data(), 32 * n_tail);
        if (me != 0) {
            // replay what rank 0's transcript went through (fri/mod.rs:37,41,54,58,69)
            for (size_t j = 0; j < n_tail; ++j) {
                if (j > 0) t->to_field();
                t->append(payload.data() + 32 * j, 32);
            }
            t->to_field();
            uint8_t be[32];
            H::to_bytes_be(H::from_lw(F->last.limbs), be);
            t->append(be, 32);
        }
        return S252_OK;
    }();
    dfree(ctx, block);
    return rc;
}

static void sharded_fri_free(ShardedFri* F) {
    for (s252_sharded_commit* sc : F->layers) s252_sharded_commit_destroy(sc);
    F->layers.clear();
    if (F->tail) s252_fri_destroy(F->tail);
    F->tail = nullptr;
}

extern "C" int s252_cairo_prove_sharded(s252_ctx* ctx, s252_comm* comm, const s252_cairo_trace* trace, size_t blowup, size_t n_queries,
                                        uint64_t coset_offset, uint8_t grinding_factor, size_t pipeline_groups, uint8_t** proof_out,
                                        size_t* proof_len) {
    NVTX_RANGE("s252_cairo_prove_sharded");
    if (!ctx || !comm || comm->ctx != ctx || !trace || !proof_out || !proof_len) return S252_ERR_INVALID;
    *proof_out = nullptr; *proof_len = 0;
    CU(ctx, cudaSetDevice(ctx->device));
    auto& A = s252nccl::api();
    const size_t G = (size_t)comm->world, me = (size_t)comm->rank;
    const size_t N = trace->n_rows, M = N * blowup, c_main = trace->n_cols, c_aux = s252::CAIRO_AUX_COLS;
    if (!is_pow2(N) || N < 2) FAIL(ctx, S252_ERR_INVALID, "trace length %zu is not a power of two", N);
    if (!is_pow2(blowup) || blowup > MAX_COSETS || coset_offset == 0) FAIL(ctx, S252_ERR_INVALID, "bad proof options");
    if (M % G || M / G < blowup) FAIL(ctx, S252_ERR_INVALID, "too many ranks for this trace");
    if (c_main < G || c_aux < G) FAIL(ctx, S252_ERR_INVALID, "fewer columns than ranks");
    const size_t rows_per = M / G, row0 = me * rows_per;
    if (pipeline_groups == 0) pipeline_groups = N >= ((size_t)1 << 20) ? 2 : 1;
    s252_cairo_trace_pin(trace);

    s252_transcript t;                                              // round_0_transcript_initialization
    s252_sharded_commit *sc_main = nullptr, *sc_aux = nullptr, *sc_comp = nullptr;
    s252_commit* comp = nullptr;
    ShardedFri fri;
    StageTimer ST;
    auto body = [&]() -> int {
        // group widths of a rank's shard: the same rule on every rank (the core exchanges them anyway)
        auto groups_of = [&](size_t n_cols_total, std::vector<size_t>* widths, size_t* lo_out) {
            size_t lo, hi, min_shard = n_cols_total;
            for (size_t r = 0; r < G; ++r) { size_t a, b; shard_range(n_cols_total, G, r, &a, &b); min_shard = std::min(min_shard, b - a); }
            shard_range(n_cols_total, G, me, &lo, &hi);
            const size_t ng = std::max<size_t>(1, std::min<size_t>(std::min<size_t>(G > 1 ? pipeline_groups : 1, min_shard), S252_MAX_PIPELINE_GROUPS));
            widths->clear();
            for (size_t g = 0; g < ng; ++g) { size_t a, b; shard_range(hi - lo, ng, g, &a, &b); widths->push_back(b - a); }
            *lo_out = lo;
        };
        // ---- round 1 (prover.rs:186-224)
        uint8_t r_main[32], r_aux[32], r_comp[32];
        std::vector<size_t> wm, wa;
        size_t lo_m, lo_a;
        groups_of(c_main, &wm, &lo_m);
        {
            size_t col = lo_m;
            auto producer = [&](size_t g, s252_commit** h) -> int {
                const int rc = commit_from_host_columns(ctx, trace->cols.data() + col * N, N, (unsigned)wm[g], blowup, coset_offset, true, h, nullptr, false);
                col += wm[g];
                return rc;
            };
            TRY(sharded_commit_core(ctx, comm, wm.data(), wm.size(), M, c_main, producer, &sc_main, r_main));
        }
        t.append(r_main, 32);
        ST.mark("main_commit");
        fe rap[3];
        for (int k = 0; k < 3; ++k) rap[k] = t.to_field();
        // build_auxiliary_trace reads main columns 19..29 (pc .. off_op1): each run is broadcast by the rank that holds it
        fe* aux = nullptr;
        {
            Tmp<fe> aux_in(ctx);
            TRY(dalloc(ctx, &aux_in.p, 11 * N));
            const size_t c0 = s252::CAIRO_PC, c1 = s252::CAIRO_PC + 11;
            for (size_t owner = 0; owner < G; ++owner) {
                size_t lo, hi;
                shard_range(c_main, G, owner, &lo, &hi);
                const size_t a = std::max(lo, c0), z = std::min(hi, c1);
                if (a >= z) continue;
                if (owner == me) {
                    size_t col = lo_m;
                    for (size_t g = 0; g < sc_main->local.size(); ++g) {          // this rank's columns, one handle per pipeline group
                        const size_t x = std::max(a, col), y = std::min(z, col + wm[g]);
                        if (x < y)
                            CU(ctx, cudaMemcpyAsync(aux_in.p + (x - c0) * N, sc_main->local[g]->trace + (x - col) * N, (y - x) * N * sizeof(fe),
                                                    cudaMemcpyDeviceToDevice, ctx->stream));
                        col += wm[g];
                    }
                }
                if (G > 1)
                    NCCL_TRY(ctx, A.Broadcast(aux_in.p + (a - c0) * N, aux_in.p + (a - c0) * N, (z - a) * N * sizeof(fe), s252nccl::kUint8, (int)owner,
                                              comm->comm, ctx->stream));
            }
            ST.mark("aux_inputs");
            int rc = dalloc(ctx, &aux, c_aux * N);
            if (rc == S252_OK) rc = cairo_build_aux(ctx, aux_in.p, s252::CAIRO_PC, N, trace->pi, rap, aux);
            rc = sharded_all_ok(ctx, comm, rc);
            if (rc != S252_OK) { dfree(ctx, aux); return rc; }
        }
        ST.mark("aux_build");
        groups_of(c_aux, &wa, &lo_a);
        {
            size_t col = lo_a;
            auto producer = [&](size_t g, s252_commit** h) -> int {
                const int rc = s252_lde_device_columns(ctx, aux + col * N, N, wa[g], blowup, coset_offset, h);
                col += wa[g];
                return rc;
            };
            const int rc = sharded_commit_core(ctx, comm, wa.data(), wa.size(), M, c_aux, producer, &sc_aux, r_aux);
            dfree(ctx, aux);
            aux = nullptr;
            TRY(rc);
        }
        t.append(r_aux, 32);
        ST.mark("aux_commit");
        // ---- round 2 (prover.rs:598-640, 226-283)
        const int nt = c_main > CA::MAIN_COLS ? 50 : 49;
        fe ba[8], bb[8], ta[50], tb[50];
        for (int k = 0; k < 8; ++k) ba[k] = t.to_field();
        for (int k = 0; k < 8; ++k) bb[k] = t.to_field();
        for (int k = 0; k < nt; ++k) ta[k] = t.to_field();
        for (int k = 0; k < nt; ++k) tb[k] = t.to_field();
        const fe *mblock = sc_main->block_cols, *ablock = sc_aux->block_cols;       // [cols][rows_per]
        {
            // the frame of the last `blowup` rows of a block reaches into the next rank's block: its first rows come over as a halo
            Tmp<fe> mhalo(ctx), ahalo(ctx), msend(ctx), asend(ctx), evals(ctx);
            TRY(dalloc(ctx, &mhalo.p, c_main * blowup));
            TRY(dalloc(ctx, &ahalo.p, c_aux * blowup));
            fe *mdst = G == 1 ? mhalo.p : nullptr, *adst = G == 1 ? ahalo.p : nullptr;
            if (G > 1) {
                TRY(dalloc(ctx, &msend.p, c_main * blowup));
                TRY(dalloc(ctx, &asend.p, c_aux * blowup));
                mdst = msend.p; adst = asend.p;
            }
            CU(ctx, cudaMemcpy2DAsync(mdst, blowup * sizeof(fe), mblock, rows_per * sizeof(fe), blowup * sizeof(fe), c_main, cudaMemcpyDeviceToDevice, ctx->stream));
            CU(ctx, cudaMemcpy2DAsync(adst, blowup * sizeof(fe), ablock, rows_per * sizeof(fe), blowup * sizeof(fe), c_aux, cudaMemcpyDeviceToDevice, ctx->stream));
            if (G > 1) {
                const int nxt = (int)((me + 1) % G), prv = (int)((me + G - 1) % G);
                NCCL_TRY(ctx, A.GroupStart());
                NCCL_TRY(ctx, A.Send(msend.p, c_main * blowup * sizeof(fe), s252nccl::kUint8, prv, comm->comm, ctx->stream));
                NCCL_TRY(ctx, A.Send(asend.p, c_aux * blowup * sizeof(fe), s252nccl::kUint8, prv, comm->comm, ctx->stream));
                NCCL_TRY(ctx, A.Recv(mhalo.p, c_main * blowup * sizeof(fe), s252nccl::kUint8, nxt, comm->comm, ctx->stream));
                NCCL_TRY(ctx, A.Recv(ahalo.p, c_aux * blowup * sizeof(fe), s252nccl::kUint8, nxt, comm->comm, ctx->stream));
                NCCL_TRY(ctx, A.GroupEnd());
            }
            TRY(dalloc(ctx, &evals.p, M));
            const CairoRowBlock B{mblock, ablock, mhalo.p, ahalo.p, rows_per, blowup, row0, rows_per};
            TRY(sharded_all_ok(ctx, comm, cairo_eval_constraints_rows(ctx, trace, B, rap, ba, bb, ta, tb, blowup, coset_offset, evals.p + row0)));
            ST.mark("constraints");
            if (G > 1) NCCL_TRY(ctx, A.AllGather(evals.p + row0, evals.p, rows_per * sizeof(fe), s252nccl::kUint8, comm->comm, ctx->stream));
            // H from its evaluations and the LDE of (H1, H2) on every rank; an H above its degree bound (the trace does not satisfy
            // the AIR) is detected here, on every rank alike
            TRY(sharded_all_ok(ctx, comm, s252_cairo_composition_lde(ctx, evals.p, N, blowup, coset_offset, &comp)));
        }
        {
            fe* cblock = nullptr;
            TRY(dalloc(ctx, &cblock, 2 * rows_per));
            CU(ctx, cudaMemcpy2DAsync(cblock, rows_per * sizeof(fe), comp->lde + row0, M * sizeof(fe), rows_per * sizeof(fe), 2, cudaMemcpyDeviceToDevice, ctx->stream));
            TRY(sharded_commit_row_block(ctx, comm, cblock, 2, M, &sc_comp, r_comp));
        }
        t.append(r_comp, 32);                                                       // prover.rs:635
        ST.mark("composition");
        // ---- round 3 (prover.rs:650-690)
        fe g;
        H::primitive_root(ilog2(N), &g);
        const fe hinv = H::inv(H::from_u64(coset_offset));
        fe z;
        for (;;) {                                                                  // sample_z_ood (transcript.rs:53-70)
            z = t.to_field();
            if (!H::eq(H::pow_u64(H::mul(z, hinv), M), H::one()) && !H::eq(H::pow_u64(z, N), H::one())) break;
        }
        const size_t cols = c_main + c_aux;
        s252_fe pts[2], z2, zlw, hz[2];
        H::to_lw(z, pts[0].limbs); H::to_lw(H::mul(z, g), pts[1].limbs); H::to_lw(H::sqr(z), z2.limbs);
        zlw = pts[0];
        std::vector<s252_fe> ood(2 * cols, s252_fe{});                              // the frame, row-major; only my columns are filled here
        {
            size_t col = lo_m;
            for (size_t gi = 0; gi < sc_main->local.size(); ++gi) { TRY(s252_commit_evaluate_at(sc_main->local[gi], pts, 2, ood.data(), cols, col)); col += wm[gi]; }
            col = c_main + lo_a;
            for (size_t gi = 0; gi < sc_aux->local.size(); ++gi) { TRY(s252_commit_evaluate_at(sc_aux->local[gi], pts, 2, ood.data(), cols, col)); col += wa[gi]; }
        }
        TRY(s252_commit_evaluate_at(comp, &z2, 1, hz, 2, 0));                       // every rank holds H1, H2
        if (G > 1) {
            Tmp<uint8_t> d(ctx);
            TRY(dalloc(ctx, &d.p, ood.size() * 32));
            CU(ctx, cudaMemcpyAsync(d.p, ood.data(), ood.size() * 32, cudaMemcpyHostToDevice, ctx->stream));
            NCCL_TRY(ctx, A.AllReduce(d.p, d.p, ood.size() * 32, s252nccl::kUint8, s252nccl::kSum, comm->comm, ctx->stream));
            CU(ctx, cudaMemcpyAsync(ood.data(), d.p, ood.size() * 32, cudaMemcpyDeviceToHost, ctx->stream));
            CU(ctx, cudaStreamSynchronize(ctx->stream));
        }
        uint8_t be[32];
        for (int k = 0; k < 2; ++k) { H::to_bytes_be(H::from_lw(hz[k].limbs), be); t.append(be, 32); }
        for (auto& v : ood) { H::to_bytes_be(H::from_lw(v.limbs), be); t.append(be, 32); }
        ST.mark("ood");
        // ---- round 4 (prover.rs:327-404)
        s252_fe gamma, gamma_p;
        H::to_lw(t.to_field(), gamma.limbs); H::to_lw(t.to_field(), gamma_p.limbs);
        std::vector<s252_fe> tg(2 * cols);
        for (auto& v : tg) H::to_lw(t.to_field(), v.limbs);
        const size_t layers = ilog2(N);
        const uint64_t offs[2] = {0, 1};
        {
            fe* p0_block = nullptr;
            int rc = dalloc(ctx, &p0_block, rows_per);
            if (rc == S252_OK) {
                DeepTables T{};
                T.cols[0] = mblock; T.strides[0] = rows_per; T.ncols[0] = (unsigned)c_main;
                T.cols[1] = ablock; T.strides[1] = rows_per; T.ncols[1] = (unsigned)c_aux;
                T.cols[2] = comp->lde + row0; T.strides[2] = M; T.ncols[2] = 2;
                T.ntables = 3;
                rc = deep_evaluate_rows(ctx, T, row0, rows_per, M, N, &zlw, offs, 2, ood.data(), &hz[0], &hz[1], &gamma, &gamma_p, tg.data(), coset_offset, p0_block);
            }
            rc = sharded_all_ok(ctx, comm, rc);
            if (rc != S252_OK) { dfree(ctx, p0_block); return rc; }
            ST.mark("deep");
            TRY(sharded_fri_commit_phase(ctx, comm, p0_block, M, layers, &t, coset_offset, sharded_fri_collapse_log(), &fri));
        }
        ST.mark("fri");
        // grinding (grinding.rs:40-48) split over the ranks: windows of about twice the expected position of the first hit
        uint8_t challenge[32];
        t.challenge(challenge);
        uint64_t nonce = ~0ull;
        {
            unsigned wl = 18;
            for (size_t x = G - 1; x; x >>= 1) ++wl;
            const unsigned window_log = std::min(32u, std::max(wl, (unsigned)grinding_factor + 1u));
            Tmp<uint64_t> d(ctx);
            TRY(dalloc(ctx, &d.p, 1));
            for (uint64_t base = 0;;) {
                uint64_t found = ~0ull;
                TRY(s252_grind_round(ctx, challenge, grinding_factor, base, 0, (unsigned)me, (unsigned)G, window_log, &found));
                if (G > 1) {
                    CU(ctx, cudaMemcpyAsync(d.p, &found, 8, cudaMemcpyHostToDevice, ctx->stream));
                    NCCL_TRY(ctx, A.AllReduce(d.p, d.p, 1, s252nccl::kUint64, s252nccl::kMin, comm->comm, ctx->stream));
                    CU(ctx, cudaMemcpyAsync(&found, d.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
                    CU(ctx, cudaStreamSynchronize(ctx->stream));
                }
                if (found != ~0ull) { nonce = found; break; }
                if (base + (1ull << window_log) <= base) FAIL(ctx, S252_ERR_NOT_FOUND, "nonce not found");
                base += 1ull << window_log;
            }
        }
        uint8_t nb[8];
        put_u64_be(nb, nonce);
        t.append(nb, 8);                                                            // prover.rs:385
        ST.mark("grinding");
        const size_t Q = layers ? n_queries : 0;                                    // fri_query_phase returns nothing without layers (fri/mod.rs:83)
        std::vector<uint64_t> iotas(Q);
        for (auto& i : iotas) i = t.to_usize() % M;                                 // every rank samples the same indices
        const size_t depth = ilog2(M), k0 = fri.tail_first;
        std::vector<s252_fe> ev(Q * layers, s252_fe{}), evs(Q * layers, s252_fe{});
        std::vector<uint8_t> pa(Q * layers * depth * 32, 0), pas(Q * layers * depth * 32, 0);
        std::vector<std::vector<s252_fe>> rows(3 + k0);
        std::vector<std::vector<uint8_t>> paths(3 + k0);
        if (Q) {
            // the three trace openings and every sharded FRI layer (iota mod size and its symmetric) in ONE reduction
            std::vector<s252_sharded_commit*> cs = {sc_main, sc_aux, sc_comp};
            std::vector<std::vector<uint64_t>> lists(3, iotas);
            for (size_t k = 0; k < k0; ++k) {
                const size_t size = M >> k;
                std::vector<uint64_t> l(2 * Q);
                for (size_t q = 0; q < Q; ++q) { l[q] = iotas[q] % size; l[Q + q] = (iotas[q] + size / 2) % size; }
                cs.push_back(fri.layers[k]);
                lists.push_back(std::move(l));
            }
            TRY(sharded_open_many(ctx, comm, cs.data(), lists.data(), cs.size(), rows.data(), paths.data()));
            for (size_t k = 0; k < k0; ++k) {
                const size_t dk = depth - k;
                for (size_t q = 0; q < Q; ++q) {
                    ev[q * layers + k] = rows[3 + k][q];
                    evs[q * layers + k] = rows[3 + k][Q + q];
                    std::memcpy(pa.data() + (q * layers + k) * depth * 32, paths[3 + k].data() + q * dk * 32, dk * 32);
                    std::memcpy(pas.data() + (q * layers + k) * depth * 32, paths[3 + k].data() + (Q + q) * dk * 32, dk * 32);
                }
            }
            if (me == 0 && layers > k0) {
                const size_t lt = layers - k0, dt = depth - k0;
                std::vector<s252_fe> tev(Q * lt), tevs(Q * lt);
                std::vector<uint8_t> tpa(Q * lt * dt * 32), tpas(Q * lt * dt * 32);
                TRY(s252_fri_query(fri.tail, iotas.data(), Q, tev.data(), tevs.data(), tpa.data(), tpas.data(), dt));
                for (size_t q = 0; q < Q; ++q)
                    for (size_t j = 0; j < lt; ++j) {
                        ev[q * layers + k0 + j] = tev[q * lt + j];
                        evs[q * layers + k0 + j] = tevs[q * lt + j];
                        std::memcpy(pa.data() + (q * layers + k0 + j) * depth * 32, tpa.data() + (q * lt + j) * dt * 32, dt * 32);
                        std::memcpy(pas.data() + (q * layers + k0 + j) * depth * 32, tpas.data() + (q * lt + j) * dt * 32, dt * 32);
                    }
            }
        }
        ST.mark("openings");
        if (me == 0) {
            ByteSink S;
            serialize_stark_proof(S, N, r_main, r_aux, r_comp, ood.data(), cols, hz, layers, fri.roots.data(), fri.last, Q, depth, evs.data(), ev.data(),
                                  pas.data(), pa.data(), rows[2].data(), paths[2].data(), rows[0].data(), c_main, paths[0].data(), rows[1].data(), c_aux,
                                  paths[1].data(), nonce);
            uint8_t* outp = (uint8_t*)std::malloc(S.b.size());
            if (!outp) FAIL(ctx, S252_ERR_INVALID, "out of host memory");
            std::memcpy(outp, S.b.data(), S.b.size());
            *proof_out = outp;
            *proof_len = S.b.size();
        }
        ST.mark("serialize");
        return S252_OK;
    };
    int rc;
    try {
        rc = body();
    } catch (const std::exception& e) {           // host allocation failure: nothing may unwind across the C boundary
        ctx->err = std::string("s252_cairo_prove_sharded: ") + e.what();
        rc = S252_ERR_INVALID;
    }
    cudaStreamSynchronize(comm->xstream);
    sharded_fri_free(&fri);
    commit_free(comp);
    s252_sharded_commit_destroy(sc_comp);
    s252_sharded_commit_destroy(sc_main);
    s252_sharded_commit_destroy(sc_aux);
    ST.mark("free");
    g_cairo_stages = ST.json + "}";
    return rc;
}
