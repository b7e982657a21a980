/*
 * stark252_b200.h -- C ABI of the B200-native LDE + commitment path of the lambdaworks Cairo prover.
 *
 * The reference (lambdaclass/lambdaworks_cairo_prover) has no FFI on this path: the seam is a set of
 * Rust generic functions.  Every entry point below names the reference
 * function it replaces (paths relative to the reference root); INTEGRATION.md shows the Rust
 * `extern "C"` block and the call-site changes a maintainer would make.
 *
 * Conventions
 *   - Field elements cross the boundary in the reference's in-memory format ("LW"): 4 x u64,
 *     limbs[0] MOST significant, Montgomery form (R = 2^256), fully reduced -- i.e. a Rust
 *     `&[FieldElement<Stark252PrimeField>]` can be passed as is (zero-copy on the host side).
 *   - Commitments / Merkle nodes are 32-byte Keccak-256 digests.
 *   - `mem` says where caller buffers live: S252_HOST (pageable or pinned host memory; copies are
 *     issued by the library) or S252_DEVICE (already resident in this GPU's HBM).
 *   - Every function returns S252_OK or a negative error code and never throws/panics across the
 *     boundary; s252_last_error() returns a message for the last failure on that context.
 *     The reference returns Result<_, FFTError> / Option and unwraps at the call sites
 *     (prover.rs:184,260,267; trace.rs:109; fri_commitment.rs:37; prover.rs:383-384).
 *   - Products of a commit stay resident on the GPU (the reference keeps Round1/Round2 alive until
 *     the proof is assembled, prover.rs:45-69) behind opaque handles freed by *_destroy.
 *   - A context owns one CUDA stream; calls on one context are serialised by the caller (the
 *     prover thread).  Use one context per thread for concurrent fine-grained calls (the
 *     reference calls evaluate_offset_fft from rayon workers, prover.rs:169-183).
 *   - There is no CPU fallback: without a CUDA device s252_ctx_create fails with S252_ERR_CUDA.
 */
#ifndef STARK252_B200_H
#define STARK252_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define S252_OK 0
#define S252_ERR_INVALID -1      /* bad argument (the reference's FFTError::InputError / debug_assert) */
#define S252_ERR_CUDA -2         /* CUDA runtime failure, or no device */
#define S252_ERR_NOT_FOUND -3    /* grinding: no nonce in range (reference: None -> expect panic) */
#define S252_ERR_RANGE -4        /* position out of range (reference: get_proof_by_pos -> None) */

#define S252_HOST 0
#define S252_DEVICE 1

typedef struct s252_ctx s252_ctx;
typedef struct s252_commit s252_commit;          /* coefficients + LDE columns + batched Merkle tree */
typedef struct s252_fri s252_fri;                /* all FRI layers: evaluations + Merkle trees */
typedef struct s252_transcript s252_transcript;  /* DefaultTranscript (host) */

/* LW element: FieldElement<Stark252PrimeField> as laid out by lambdaworks-math */
typedef struct { uint64_t limbs[4]; } s252_fe;

/* ---- context ---------------------------------------------------------------------------- */
int s252_ctx_create(int device, s252_ctx **out);
void s252_ctx_destroy(s252_ctx *ctx);
const char *s252_last_error(const s252_ctx *ctx);
/* Block until everything issued on the context's stream has finished. */
int s252_ctx_synchronize(s252_ctx *ctx);
/* The context's CUDA stream as a cudaStream_t (for event timing by the caller). */
void *s252_ctx_stream(s252_ctx *ctx);
/* Number of kernels this library has launched on the context since creation. */
uint64_t s252_ctx_launch_count(const s252_ctx *ctx);
/* Return the cached (currently unused) device blocks of the context's arena to the driver. */
int s252_ctx_trim(s252_ctx *ctx);
/* Per-kernel device timing: enable = 1 starts recording CUDA events around every kernel launch on
 * this context's stream, 2 also clears the accumulated totals, 0 stops.  s252_ctx_profile_read
 * synchronises and writes a JSON object {"kernel": {"launches": n, "ms": t}, ..} into buf. */
int s252_ctx_profile(s252_ctx *ctx, int enable);
int s252_ctx_profile_read(s252_ctx *ctx, char *buf, size_t cap);
/* Device allocation helpers for callers that want S252_DEVICE buffers without another runtime. */
int s252_device_alloc(s252_ctx *ctx, size_t bytes, void **out);
int s252_device_free(s252_ctx *ctx, void *ptr);
int s252_copy_to_device(s252_ctx *ctx, void *dst, const void *src, size_t bytes);
int s252_copy_to_host(s252_ctx *ctx, void *dst, const void *src, size_t bytes);
/* Page-lock a caller buffer (a Rust Vec<FE>, say) so that S252_HOST calls can stream it: with pinned memory
 * s252_interpolate_and_commit / s252_interpolate_and_lde read the table over PCIe one column group at a time while
 * the previous group is being transformed (pageable memory is copied in one piece first).  Registration costs
 * about a millisecond per 10 MB: do it once for a buffer that is reused. */
int s252_host_register(void *ptr, size_t bytes);
int s252_host_unregister(void *ptr);
/* Strided host -> device copy on the context's stream: `height` runs of `width` bytes, `src_pitch` / `dst_pitch` bytes apart
 * (asynchronous with pinned host memory).  A GPU that shares one column's transform uploads only its slab with it. */
int s252_copy_2d_to_device(s252_ctx *ctx, void *dst, size_t dst_pitch, const void *src, size_t src_pitch, size_t width, size_t height);
/* Prefetch of the NEXT trace while the current one is being committed: the copy is queued on a
 * second stream and returns at once (pinned host memory gives a true asynchronous DMA);
 * s252_copy_stream_wait makes all later work on the compute stream wait for the prefetches issued
 * so far.  dst must not be read or written by work already queued on the compute stream. */
int s252_copy_to_device_async(s252_ctx *ctx, void *dst, const void *src, size_t bytes);
int s252_copy_stream_wait(s252_ctx *ctx);

/* ---- FFTPoly: fine-grained entry points ------------------------------------------------ */
/* Polynomial::interpolate_fft(evals)                       -- call site src/starks/trace.rs:107
 * n must be a power of two; writes n coefficients (trailing zeros are NOT trimmed). */
int s252_interpolate_fft(s252_ctx *ctx, const s252_fe *evals, size_t n, s252_fe *coeffs, int mem);
/* Polynomial::interpolate_offset_fft(evals, offset)        -- src/starks/constraints/evaluation_table.rs:32 */
int s252_interpolate_offset_fft(s252_ctx *ctx, const s252_fe *evals, size_t n, const s252_fe *offset,
                                s252_fe *coeffs, int mem);
/* Length evaluate_offset_fft will produce: max(coeff_len, domain_size).next_power_of_two() * blowup,
 * with coeff_len = n_coeffs (pass the trimmed length, as Polynomial::new trims). domain_size 0 = None. */
size_t s252_evaluate_offset_fft_len(size_t n_coeffs, size_t blowup, size_t domain_size);
/* p.evaluate_offset_fft(blowup, domain_size, offset)       -- src/starks/prover.rs:117,
 * src/starks/fri/fri_commitment.rs:36.  out[i] = p(offset * w_len^i), natural order. */
int s252_evaluate_offset_fft(s252_ctx *ctx, const s252_fe *coeffs, size_t n_coeffs, size_t blowup,
                             size_t domain_size, const s252_fe *offset, s252_fe *out, size_t out_capacity, int mem);
/* evaluate_polynomial_on_lde_domain(p, blowup, domain_size, offset) -- src/starks/prover.rs:106-123
 * (incl. the step rule); writes domain_size*blowup elements. */
int s252_evaluate_polynomial_on_lde_domain(s252_ctx *ctx, const s252_fe *coeffs, size_t n_coeffs, size_t blowup,
                                           size_t domain_size, const s252_fe *offset, s252_fe *out, int mem);

/* One column's transform shared by `parts` GPUs (a single oversized column; four-step NTT with
 * one all-to-all).  The column is a 2^l1 x (N / 2^l1) matrix (first digit x inner position, natural index = k1 * inner + i
 * on the way in).  phase 0: the first pass on this GPU's inner positions [part * inner / parts, ..) of `in`, result left in z
 * at the same positions; the caller then moves rows k1 in [part * 2^l1 / parts, ..) of z to this GPU (NCCL all-to-all);
 * phase 1: the remaining passes on those rows: this GPU's outputs k = k1 + 2^l1 * q (runs of 2^l1 / parts values, times
 * n_cosets) land in `out`, natural order.  phase 2: only writes l1 to *log_l1.
 * inverse != 0: Polynomial::interpolate_fft (n_cosets = 1); else evaluate_offset_fft(n_cosets, Some(N), coset_offset).
 * in: N elements, z: n_cosets * N, out: n_cosets * N -- device, library-internal element format (s252_commit_device_lde).
 * lambdaworks_cairo_prover_b200/column_distributed.py drives it. */
int s252_ntt_shared(s252_ctx *ctx, unsigned log_n, int inverse, size_t n_cosets, uint64_t coset_offset, int phase, unsigned part,
                    unsigned parts, const void *in, void *z, void *out, unsigned *log_l1);
/* Element-format change between the reference's LW layout and the library-internal one on device buffers (n elements;
 * to_internal != 0: LW -> internal).  in == out is allowed. */
int s252_convert_elements(s252_ctx *ctx, const void *in, void *out, size_t n, int to_internal);

/* ---- round 1 / round 2 commits ------------------------------------------------------------ */
/* interpolate_and_commit(trace, domain, transcript)         -- src/starks/prover.rs:126-159
 * trace: row-major n_rows x n_cols (TraceTable.table, src/starks/trace.rs:9-13).
 * Does compute_trace_polys (trace.rs:104), compute_lde_trace_evaluations (prover.rs:161-185) and
 * batch_commit (prover.rs:96-104); the caller appends `root` to its transcript (prover.rs:151). */
int s252_interpolate_and_commit(s252_ctx *ctx, const s252_fe *trace, size_t n_rows, size_t n_cols, size_t blowup,
                                uint64_t coset_offset, int mem, s252_commit **out, uint8_t root[32]);
/* The first two steps of interpolate_and_commit only (compute_trace_polys, trace.rs:104, and
 * compute_lde_trace_evaluations, prover.rs:161-185): the handle holds coefficients and LDE columns
 * but no tree.  One rank of a column-sharded commit calls this on its columns. */
int s252_interpolate_and_lde(s252_ctx *ctx, const s252_fe *trace, size_t n_rows, size_t n_cols, size_t blowup,
                             uint64_t coset_offset, int mem, s252_commit **out);
/* batch_commit (prover.rs:96-104) over column-major columns already resident on this device in the
 * library's internal element format (cols[j*col_stride + i], as returned by
 * s252_commit_device_lde): e.g. a row block assembled from the LDE shards of several GPUs. */
int s252_commit_device_columns(s252_ctx *ctx, const void *cols, size_t col_stride, size_t n_cols, size_t n_rows,
                               s252_commit **out, uint8_t root[32]);
/* The same without the copy: the tree is built over the caller's columns (col_stride must equal n_rows),
 * which must stay alive and unchanged while the handle is in use.  root may be NULL: the call then returns without
 * waiting for the device and the root stays at s252_commit_device_nodes(handle) (a sharded commit gathers the subtree
 * roots of all GPUs device to device). */
int s252_commit_device_columns_inplace(s252_ctx *ctx, const void *cols, size_t col_stride, size_t n_cols, size_t n_rows,
                                       s252_commit **out, uint8_t root[32]);
/* ---- ONE trace committed by the GPUs of one box -----------------------------------------------
 * interpolate_and_commit (src/starks/prover.rs:126-159) with the columns of the trace sharded over G GPUs: one process or
 * thread per GPU, each with its own s252_ctx; the entry points below are COLLECTIVE (every rank calls them with the same shape
 * arguments, in the same order).  NCCL is called from inside the library (bound at run time: dlopen of libnccl.so.2, or the
 * name in S252_NCCL_LIB), so a Rust/C caller needs no Python and no torch.  Rank r owns the contiguous column range
 * [r*c/G + min(r, c%G), ..) (33 columns over 8 ranks: 5,4,4,4,4,4,4,4 -- SURVEY.md 8d config C4); after the call it holds the
 * coefficients + LDE of its columns for ALL rows and ALL columns for its block of n_rows*blowup/G LDE rows with the Merkle
 * subtree over them; the top log2(G) levels are replicated.  G must be a power of two. */
#define S252_COMM_ID_BYTES 128
#define S252_MAX_PIPELINE_GROUPS 16
typedef struct s252_comm s252_comm;                      /* one rank's end of a communicator */
typedef struct s252_sharded_commit s252_sharded_commit;  /* what one rank holds after a sharded commit */
/* Rank 0 creates an id (ncclGetUniqueId) and hands it to the other ranks by any means; every rank then joins. */
int s252_comm_unique_id(uint8_t id[S252_COMM_ID_BYTES]);
int s252_comm_create(s252_ctx *ctx, const uint8_t id[S252_COMM_ID_BYTES], int rank, int world, s252_comm **out);
void s252_comm_destroy(s252_comm *comm);
int s252_comm_rank(const s252_comm *comm);
int s252_comm_world(const s252_comm *comm);
/* group_tables[g]: row-major TraceTable [n_rows][group_cols[g]] (LW elements; `mem` = S252_HOST or S252_DEVICE) holding pipeline
 * group g of this rank's columns -- the groups, in order, are the rank's column range.  The exchange of group g (one ncclSend per
 * column and destination straight out of the LDE buffer, one ncclRecv per column and source straight into the row block) runs on
 * a second stream under the upload + transforms of group g+1.  root: the Merkle root of the whole table, the same on every
 * rank and equal to s252_interpolate_and_commit's on one GPU. */
int s252_interpolate_and_commit_sharded(s252_ctx *ctx, s252_comm *comm, const s252_fe *const *group_tables, const size_t *group_cols,
                                        size_t n_groups, size_t n_rows, size_t n_cols_total, size_t blowup, uint64_t coset_offset,
                                        int mem, s252_sharded_commit **out, uint8_t root[32]);
void s252_sharded_commit_destroy(s252_sharded_commit *sc);
size_t s252_sharded_commit_n_rows(const s252_sharded_commit *sc);   /* LDE rows of the whole table */
size_t s252_sharded_commit_n_cols(const s252_sharded_commit *sc);   /* columns of the whole table */
/* This rank's columns (coefficients + LDE over all rows), one handle per pipeline group; borrowed, freed with the sharded commit. */
size_t s252_sharded_commit_n_local(const s252_sharded_commit *sc);
s252_commit *s252_sharded_commit_local(const s252_sharded_commit *sc, size_t group);
/* All columns over this rank's row block + the subtree (borrowed). */
s252_commit *s252_sharded_commit_block(const s252_sharded_commit *sc);
/* MerkleTree::get_proof_by_pos + the opened rows for global positions (open_deep_composition_poly, prover.rs:484-529), on every
 * rank: rows_out [n_idx][n_cols] LW, paths_out [n_idx][log2(n_rows)][32] leaf -> root.  The owner of a row serves its values and
 * the subtree part of the path (one all-reduce of a packed buffer), the top levels are replicated.  S252_ERR_RANGE if a position
 * is out of range (reference: None). */
int s252_sharded_commit_open(s252_sharded_commit *sc, const uint64_t *indices, size_t n_idx, s252_fe *rows_out, uint8_t *paths_out);

/* Round 2 (src/starks/prover.rs:254-276): evaluate_polynomial_on_lde_domain for each of n_polys
 * polynomials (polys: n_polys x n_coeffs, polynomial-major; n_coeffs <= domain_size) and
 * batch_commit over the zipped rows. */
int s252_lde_and_commit(s252_ctx *ctx, const s252_fe *polys, size_t n_coeffs, size_t n_polys, size_t domain_size,
                        size_t blowup, uint64_t coset_offset, int mem, s252_commit **out, uint8_t root[32]);
/* BatchedMerkleTree::build(rows) / FriMerkleTree::build(evals) -- src/starks/prover.rs:101,
 * src/starks/fri/fri_commitment.rs:39.  rows: row-major n_rows x n_cols; n_rows a power of two. */
int s252_merkle_build(s252_ctx *ctx, const s252_fe *rows, size_t n_rows, size_t n_cols, int mem, s252_commit **out,
                      uint8_t root[32]);
void s252_commit_destroy(s252_commit *c);
size_t s252_commit_n_cols(const s252_commit *c);
size_t s252_commit_n_rows(const s252_commit *c);       /* rows of the committed (LDE) table */
size_t s252_commit_n_coeffs(const s252_commit *c);     /* coefficients kept per column (0 if none) */
/* tree.root */
int s252_commit_root(const s252_commit *c, uint8_t root[32]);
/* Copy out LDE column `col`, rows [first, first+count)       -- reads of lde_trace, prover.rs:243 */
int s252_commit_read_lde(s252_commit *c, size_t col, size_t first, size_t count, s252_fe *out);
/* Copy out the coefficients of column `col`                 -- reads of trace_polys, prover.rs:314,360 */
int s252_commit_read_coeffs(s252_commit *c, size_t col, s252_fe *out);
/* Copy out Merkle nodes [first, first+count) of the heap array (root = node 0, leaf i = n-1+i). */
int s252_commit_read_nodes(s252_commit *c, size_t first, size_t count, uint8_t *out);
/* open_deep_composition_poly's reads (src/starks/prover.rs:484-529): for each of n_idx row indices,
 * rows_out[q*n_cols + j] = LDE value and paths_out[(q*depth + k)*32..] = tree.get_proof_by_pos(idx)
 * merkle_path entry k (leaf -> root), depth = log2(n_rows).  Either output may be NULL. */
int s252_commit_open(s252_commit *c, const uint64_t *indices, size_t n_idx, s252_fe *rows_out, uint8_t *paths_out);
/* Raw device pointers (column-major, library-internal element format) for fusing
 * later prover stages on the GPU. */
const void *s252_commit_device_lde(const s252_commit *c);
const void *s252_commit_device_coeffs(const s252_commit *c);
const void *s252_commit_device_nodes(const s252_commit *c);

/* ---- FRI ------------------------------------------------------------------------------------- */
/* fri_commit_phase(number_layers, p_0, transcript, coset_offset, domain_size)
 *                                                           -- src/starks/fri/mod.rs:20-72
 * p0: n_coeffs coefficients (n_coeffs <= domain_size).  Appends every layer root, samples every
 * zeta and appends the last value through `transcript`, exactly in the reference's order.
 * roots_out: number_layers x 32 bytes (may be NULL). */
int s252_fri_commit_phase(s252_ctx *ctx, size_t number_layers, const s252_fe *p0, size_t n_coeffs,
                          s252_transcript *transcript, const s252_fe *coset_offset, size_t domain_size, int mem,
                          s252_fri **out, s252_fe *last_value, uint8_t *roots_out);
/* The same phase layer by layer, for a caller that keeps its own (Rust) transcript between the calls:
 *   s252_fri_layer0       FriLayer::new(p0, coset_offset, domain_size)          -- fri/mod.rs:33-35, fri_commitment.rs:30-47
 *   s252_fri_fold_commit  fold_polynomial(zeta) + FriLayer::new of the next layer -- fri/mod.rs:43-51; returns its root
 *   s252_fri_fold_last    the last fold and fri_last_value                      -- fri/mod.rs:58-66
 * The handle grows by one layer per s252_fri_fold_commit and serves s252_fri_query / s252_fri_read_* as usual. */
int s252_fri_layer0(s252_ctx *ctx, const s252_fe *p0, size_t n_coeffs, const s252_fe *coset_offset, size_t domain_size, int mem,
                    s252_fri **out, uint8_t root[32]);
int s252_fri_fold_commit(s252_fri *f, const s252_fe *zeta, uint8_t root[32]);
int s252_fri_fold_last(s252_fri *f, const s252_fe *zeta, s252_fe *last_value);
/* Round 3 reads: Frame::get_trace_evaluations (src/starks/frame.rs:67-83) and
 * H1(z^2), H2(z^2) (prover.rs:296-300) from the coefficients resident in a commit:
 * out[p*out_stride + col_offset + j] = poly_j(points[p]).  Several commits (main, aux) fill one
 * frame by using different col_offset. */
int s252_commit_evaluate_at(s252_commit *c, const s252_fe *points, size_t n_points, s252_fe *out, size_t out_stride,
                            size_t col_offset);
/* Round 4 from the resident commits (src/starks/prover.rs:327-404 minus the challenge sampling, which
 * the caller does): builds the DEEP composition polynomial (compute_deep_composition_poly,
 * prover.rs:410-482) directly as evaluations on the LDE coset,
 *   p0(x) = sum_k [sum_j gammas[j*K+k] (t_j(x) - trace_ood[k*cols+j])] / (x - z g^offset_k)
 *         + [gamma (H1(x) - h1_z2) + gamma_p (H2(x) - h2_z2)] / (x - z^2),
 * and runs fri_commit_phase on it.  trace_commits: the round-1 commits in column order (main, aux);
 * trace_ood: K x total_cols (Frame data, row-major); trace_gammas: total_cols x K in the
 * reference's sampling order (prover.rs:352-355, index i*K + k). */
int s252_fri_commit_phase_deep(s252_ctx *ctx, size_t number_layers, s252_commit *const *trace_commits,
                               size_t n_trace_commits, s252_commit *composition_commit, const s252_fe *z,
                               const uint64_t *transition_offsets, size_t n_offsets, const s252_fe *trace_ood,
                               const s252_fe *h1_z2, const s252_fe *h2_z2, const s252_fe *gamma, const s252_fe *gamma_p,
                               const s252_fe *trace_gammas, s252_transcript *transcript, uint64_t coset_offset,
                               s252_fri **out, s252_fe *last_value, uint8_t *roots_out);
/* fri_commit_phase from layer 0 given as evaluations on the LDE coset (device, internal element format):
 * what s252_fri_commit_phase_deep runs after building the DEEP polynomial; a sharded prover gathers the
 * row blocks of that polynomial and calls this on one rank. */
int s252_fri_commit_phase_evals(s252_ctx *ctx, size_t number_layers, const void *p0_evals, size_t domain_size,
                                s252_transcript *transcript, uint64_t coset_offset, s252_fri **out, s252_fe *last_value,
                                uint8_t *roots_out);
/* Building blocks of a commit phase sharded over several GPUs by row blocks (SURVEY.md section 8e, row 5;
 * lambdaworks_cairo_prover_b200/fri_distributed.py): one fold of fri/mod.rs:43-51 on the rows [i0, i0+count) of the next
 * layer with its operands given separately (v[j] = layer_k[i0+j], s[j] = layer_k[i0+j+layer_size/2]; they arrive from
 * two other GPUs), and the rest of the phase on one GPU from layer `layer_index` on, given in full. */
int s252_fri_fold_rows(s252_ctx *ctx, const void *v, const void *s, size_t count, size_t i0, size_t layer_size,
                       size_t domain_size, size_t layer_index, const s252_fe *zeta, uint64_t coset_offset, void *out);
int s252_fri_commit_phase_from_layer(s252_ctx *ctx, size_t number_layers, const void *evals, size_t layer_size,
                                     s252_transcript *transcript, uint64_t coset_offset, size_t layer_index, s252_fri **out,
                                     s252_fe *last_value, uint8_t *roots_out);
void s252_fri_destroy(s252_fri *f);
size_t s252_fri_n_layers(const s252_fri *f);
/* FriLayer.evaluation[first..first+count) of layer k */
int s252_fri_read_layer(s252_fri *f, size_t layer, size_t first, size_t count, s252_fe *out);
int s252_fri_read_nodes(s252_fri *f, size_t layer, size_t first, size_t count, uint8_t *out);
/* fri_query_phase's reads (src/starks/fri/mod.rs:74-127) for the given iotas: per query q and layer k
 * (size_k = domain_size >> k): evals[q*L+k] = evaluation[iota % size_k],
 * evals_sym[q*L+k] = evaluation[(iota + size_k/2) % size_k] and their auth paths, each padded to
 * `path_stride` digests (use log2(domain_size)); path k has log2(size_k) entries. */
int s252_fri_query(s252_fri *f, const uint64_t *iotas, size_t n_queries, s252_fe *evals, s252_fe *evals_sym,
                   uint8_t *paths, uint8_t *paths_sym, size_t path_stride);

/* ---- grinding --------------------------------------------------------------------------------- */
/* generate_nonce_with_grinding(challenge, grinding_factor)  -- src/starks/grinding.rs:40-48
 * Returns the SMALLEST nonce; S252_ERR_NOT_FOUND if none below `limit` (0 = 2^64-1). */
int s252_generate_nonce_with_grinding(s252_ctx *ctx, const uint8_t challenge[32], uint8_t grinding_factor,
                                      uint64_t limit, uint64_t *nonce);

/* The same search shared by `parts` GPUs (one process per GPU): one round over the window [base, base + 2^window_log)
 * (18 <= window_log <= 40), of which this GPU tests batches part, part + parts, .. of 2^18 nonces.  *found = this GPU's
 * smallest accepted nonce or UINT64_MAX.  The caller takes the MIN over the GPUs (an all-reduce) -- that is the reference's
 * nonce -- and calls again with the next window if no GPU found one.  Use window_log ~ grinding_factor + 1: a GPU does not
 * see the other GPUs' hits while its kernel runs. */
int s252_grind_round(s252_ctx *ctx, const uint8_t challenge[32], uint8_t grinding_factor, uint64_t base, uint64_t limit,
                     unsigned part, unsigned parts, unsigned window_log, uint64_t *found);

/* ByteConversion::to_bytes_be for n elements on the host (the proof's wire format): out = n x 32 bytes. */
void s252_fe_to_bytes_be(const s252_fe *in, size_t n, uint8_t *out);
/* Keccak256 on the host (node rule of the Merkle back-ends: Keccak256(left || right)). */
void s252_keccak256(const uint8_t *data, size_t len, uint8_t out[32]);

/* ---- transcript (host) -------------------------------------------------------------------------- */
/* DefaultTranscript (lambdaworks-crypto) and the helpers of src/starks/transcript.rs:13-51 */
s252_transcript *s252_transcript_new(void);
void s252_transcript_free(s252_transcript *t);
void s252_transcript_append(s252_transcript *t, const uint8_t *data, size_t len);
void s252_transcript_challenge(s252_transcript *t, uint8_t out[32]);
void s252_transcript_to_field(s252_transcript *t, s252_fe *out);
uint64_t s252_transcript_to_usize(s252_transcript *t);

/* ---- diagnostics ----------------------------------------------------------------------------- */
/* Integer-pipe peak micro-benchmarks on this device (the roofline denominators that are not in
 * MEASURED_PEAKS.json): results in Gops/s of 32-bit lane-operations.
 * out[0] = IMAD.WIDE.U32 (pure 32x32->64 products, data-dependent multiplicands), out[1] = LOP3, out[2] = SHF (funnel shift),
 * out[3] = IADD3 with carry, out[4] = IMAD.WIDE and LOP3 interleaved 1:1 (sum of both), out[5] = wide multiply-adds in the
 * field multiply's carry rows (IMAD.WIDE.U32.X), out[6] = IMAD (32-bit mad.lo), out[7] = IMAD.WIDE.U32 with both
 * multiplicands in registers of equal parity. */
int s252_microbench_int_pipes(s252_ctx *ctx, double out[8]);
/* Montgomery multiplications per second (Gmul/s) of fe_mul in a register-resident loop. */
int s252_microbench_fe_mul(s252_ctx *ctx, double *gmuls);
/* Keccak-f[1600] permutations per second (Gperm/s), register-resident. */
int s252_microbench_keccak(s252_ctx *ctx, double *gperms);
/* Element-wise field ops on device buffers of LW elements, for parity tests:
 * op 0: a*b, 1: a+b, 2: a-b, 3: a^-1 (b ignored). */
int s252_fe_binop(s252_ctx *ctx, int op, const s252_fe *a, const s252_fe *b, s252_fe *out, size_t n, int mem);
/* Keccak256 of n messages of msg_len bytes each (host buffers), through the device Keccak. */
int s252_keccak256_batch(s252_ctx *ctx, const uint8_t *msgs, size_t msg_len, size_t n, uint8_t *digests);

#ifdef __cplusplus
}
#endif
#endif
