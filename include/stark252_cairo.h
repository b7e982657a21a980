/*
 * stark252_cairo.h -- C ABI of the Cairo side of the B200-native prover: the callers and data formats
 * on either side of the LDE + commitment path, in the same shared library as
 * include/stark252_b200.h.
 *
 *   front-end (host)   a minimal Cairo-0 machine standing in for cairo-vm as the reference drives it
 *                      (src/cairo/runner/run.rs:62-241) and build_main_trace
 *                      (src/cairo/execution_trace.rs:57-87);
 *   prover (GPU)       generate_cairo_proof (src/cairo/air.rs:1183-1190) = prove::<CairoAIR>
 *                      (src/starks/prover.rs:532-776): both round-1 commits, the auxiliary (RAP) trace,
 *                      the 49 Cairo transition constraints + 8 boundary constraints evaluated over the
 *                      LDE coset on the device, composition polynomial, out-of-domain frame, DEEP
 *                      composition polynomial, FRI, grinding, openings, and
 *                      StarkProof::serialize (src/starks/proof/stark.rs:161-218).
 *
 * Binary formats are the reference's own: register trace = 24-byte little-endian rows (ap, fp, pc),
 * src/cairo/register_states.rs:47-78; memory = 40-byte rows (u64 LE address, 32-byte LE value),
 * src/cairo/cairo_mem.rs:35-61; field elements in tables = LW layout (see stark252_b200.h).
 * Functions without a context report failures through s252_cairo_last_error() (thread local).
 */
#ifndef STARK252_CAIRO_H
#define STARK252_CAIRO_H
#include "stark252_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct s252_cairo_run s252_cairo_run;       /* relocated register trace + memory of one execution */
typedef struct s252_cairo_trace s252_cairo_trace;   /* main trace table + PublicInputs */

/* PublicInputs (src/cairo/air.rs:155-176) without the public memory map (read it with
 * s252_cairo_trace_public_memory). */
typedef struct {
    uint64_t pc_init, ap_init, fp_init, pc_final, ap_final;
    uint64_t num_steps;
    uint64_t n_public_memory;
    uint64_t rc_segment[2];        /* MemorySegment::RangeCheck range, valid if has_rc_segment */
    uint64_t output_segment[2];    /* MemorySegment::Output range, valid if has_output_segment */
    uint16_t range_check_min, range_check_max;   /* valid if has_range_check_bounds */
    uint8_t has_range_check_bounds, has_rc_segment, has_output_segment, reserved;
} s252_cairo_public_inputs;

const char *s252_cairo_last_error(void);

/* ---- front-end ---------------------------------------------------------------------------- */
/* Runs a hint-free, builtin-free Cairo-0 program the way run_program(None, layout, .., V0) does
 * (run.rs:62-100: proof_mode = false, relocation on): program words (32-byte big-endian each, the
 * "data" array of the compiled JSON) are loaded at address 1, `main` is entered at
 * 1 + entry_offset with the stack [return_fp, end], and execution stops at the final `ret`. */
int s252_cairo_vm_run(const uint8_t *program_be, size_t n_words, uint64_t entry_offset, uint64_t max_steps,
                      s252_cairo_run **out);
/* The same for programs that declare builtins (cairo-vm's BuiltinRunner for the two the AIR knows, src/cairo/air.rs:594-625):
 * builtins = S252_BUILTIN_OUTPUT | S252_BUILTIN_RANGE_CHECK.  main receives the segment base pointers on the stack below
 * [return_fp, end], output first (the %builtins order); the segments are relocated behind the execution segment.
 * s252_cairo_run_segment reports the relocated [begin, end) of a segment (which: 0 = range_check, 1 = output) -- what
 * generate_prover_args passes on as range_check_builtin_range / output range (run.rs:243-266); returns 1 if present. */
#define S252_BUILTIN_OUTPUT 1u
#define S252_BUILTIN_RANGE_CHECK 2u
int s252_cairo_vm_run_builtins(const uint8_t *program_be, size_t n_words, uint64_t entry_offset, uint64_t max_steps,
                               unsigned builtins, s252_cairo_run **out);
int s252_cairo_run_segment(const s252_cairo_run *run, int which, uint64_t range[2]);
void s252_cairo_run_destroy(s252_cairo_run *run);
size_t s252_cairo_run_steps(const s252_cairo_run *run);
size_t s252_cairo_run_trace_len(const s252_cairo_run *run);    /* bytes: 24 per step */
size_t s252_cairo_run_memory_len(const s252_cairo_run *run);   /* bytes: 40 per cell */
void s252_cairo_run_trace_bytes(const s252_cairo_run *run, uint8_t *out);
void s252_cairo_run_memory_bytes(const s252_cairo_run *run, uint8_t *out);

/* PublicInputs::from_regs_and_mem (air.rs:183-214) + build_main_trace (execution_trace.rs:57-87):
 * execution trace, range-check holes, memory holes, public-memory dummy accesses, power-of-two
 * padding.  rc_range / output_range: [start, end) or NULL (generate_prover_args, run.rs:243-266). */
int s252_cairo_build_main_trace(const uint8_t *trace_le, size_t trace_len, const uint8_t *memory_le, size_t memory_len,
                                size_t program_size, const uint64_t *rc_range, const uint64_t *output_range,
                                s252_cairo_trace **out);
/* The execution rows only: build_cairo_execution_trace (execution_trace.rs:261-356). */
int s252_cairo_build_execution_trace(const uint8_t *trace_le, size_t trace_len, const uint8_t *memory_le,
                                     size_t memory_len, size_t program_size, const uint64_t *rc_range,
                                     const uint64_t *output_range, s252_cairo_trace **out);
void s252_cairo_trace_destroy(s252_cairo_trace *t);
/* Registers the table's pages with the CUDA driver so that round 1 uploads it by DMA (done implicitly
 * by the first s252_cairo_round1 / s252_cairo_prove on the handle; call it earlier to keep the
 * one-off registration cost out of the first proof). */
int s252_cairo_trace_pin(const s252_cairo_trace *t);
size_t s252_cairo_trace_n_rows(const s252_cairo_trace *t);
size_t s252_cairo_trace_n_cols(const s252_cairo_trace *t);
const s252_fe *s252_cairo_trace_table(const s252_cairo_trace *t);   /* row-major n_rows x n_cols, LW */
void s252_cairo_trace_public_inputs(const s252_cairo_trace *t, s252_cairo_public_inputs *out);
/* addrs[n_public_memory], values[n_public_memory] (LW), sorted by address */
void s252_cairo_trace_public_memory(const s252_cairo_trace *t, uint64_t *addrs, s252_fe *values);
/* PublicInputs::serialize (air.rs:217-276); public memory in address order (the reference iterates a
 * HashMap, so its order is not reproducible).  Returns the length; writes if out != NULL. */
size_t s252_cairo_trace_serialize_public_inputs(const s252_cairo_trace *t, uint8_t *out);

/* A trace handle around a caller-built table (a Rust TraceTable + PublicInputs): table row-major LW
 * (copied), public memory as parallel arrays.  n_rows must be a power of two.
 * LIMIT: build_auxiliary_trace on the device sorts the four address columns by their low 64 bits and the three
 * offset columns by their low 16 bits (machine words: what build_main_trace produces); the reference sorts full
 * representatives (air.rs:529-533,684-689), so a hand-made table with wider values in those columns is outside
 * the supported range. */
int s252_cairo_trace_from_table(const s252_fe *table, size_t n_rows, size_t n_cols, const s252_cairo_public_inputs *pub,
                                const uint64_t *pub_addrs, const s252_fe *pub_values, s252_cairo_trace **out);

/* ---- prover (GPU) ------------------------------------------------------------------------- */
/* Round 1 of prove::<CairoAIR> (round_1_randomized_air_with_preprocessing, src/starks/prover.rs:186-224):
 * interpolate_and_commit(main) -> transcript.append(root) -> build_rap_challenges (air.rs:731-737) ->
 * build_auxiliary_trace ON THE DEVICE (air.rs:660-729: stable radix sort by address, permutation-argument
 * columns as multiplicative scans) -> interpolate_and_commit(aux) -> transcript.append(root).
 * rap_out[3] = alpha_memory, z_memory, z_range_check. */
int s252_cairo_round1(s252_ctx *ctx, const s252_cairo_trace *trace, size_t blowup, uint64_t coset_offset,
                      s252_transcript *transcript, s252_commit **main_out, s252_commit **aux_out, s252_fe rap_out[3]);
/* Trace evaluations kept in a round-1 handle (column `col`, n_coeffs values): reads the device-built
 * auxiliary trace back for parity tests. */
int s252_commit_read_trace(s252_commit *c, size_t col, s252_fe *out);
/* Round 2 (prover.rs:598-640 + round_2_compute_composition_polynomial :226-283): samples the boundary
 * and transition coefficients from the transcript, evaluates the 49 (50 with the range-check builtin)
 * transition constraints and the 8 boundary constraints of CairoAIR over the LDE coset on the device
 * (ConstraintEvaluator::evaluate, constraints/evaluator.rs:40-262), interpolates H
 * (interpolate_offset_fft), splits it into H1/H2, extends and commits them, appends the root.
 * The handle keeps the H1, H2 coefficients (s252_commit_read_coeffs).  S252_ERR_INVALID if H exceeds
 * its degree bound 2N (the trace does not satisfy the AIR). */
int s252_cairo_round2(s252_ctx *ctx, const s252_cairo_trace *trace, s252_commit *main_commit, s252_commit *aux_commit,
                      const s252_fe rap[3], size_t blowup, uint64_t coset_offset, s252_transcript *transcript,
                      s252_commit **composition_out);
/* ConstraintEvaluator::evaluate alone (constraints/evaluator.rs:40-262) with caller-supplied coefficients:
 * boundary_coeffs = 8 (alpha, beta) pairs, transition_coeffs = 49 (50 with the range-check builtin)
 * pairs; out = n_rows*blowup evaluations of the composition polynomial on the LDE coset (LW, host).
 * Works for any table (the AIR need not be satisfied): the parity tests drive it with random traces. */
int s252_cairo_constraint_evaluations(s252_ctx *ctx, const s252_cairo_trace *trace, s252_commit *main_commit,
                                      s252_commit *aux_commit, const s252_fe rap[3], const s252_fe *boundary_coeffs,
                                      const s252_fe *transition_coeffs, size_t blowup, uint64_t coset_offset, s252_fe *out);
/* generate_cairo_proof (src/cairo/air.rs:1183-1190) = prove::<Stark252PrimeField, CairoAIR>: rounds 1-4
 * on the device and StarkProof::serialize (src/starks/proof/stark.rs:161-218).  *proof_out is malloc'ed
 * by the library; release it with s252_cairo_proof_free. */
int s252_cairo_prove(s252_ctx *ctx, const s252_cairo_trace *trace, size_t blowup, size_t fri_number_of_queries,
                     uint64_t coset_offset, uint8_t grinding_factor, uint8_t **proof_out, size_t *proof_len);
void s252_cairo_proof_free(uint8_t *proof);
/* The same proof as ONE collective call over the GPUs of one box: every rank -- one process or thread per GPU, each
 * with its own context and its end of an s252_comm (stark252_b200.h) -- calls it with the same trace and options.  Columns are
 * sharded for the LDE, rows for the trees, the constraint evaluation, the DEEP polynomial, FRI (pairwise fold exchange, collapse
 * to rank 0 at 2^19 evaluations) and the openings; grinding is split over the ranks; NCCL is called from the library.  Rank 0
 * receives StarkProof::serialize bytes, byte-identical to s252_cairo_prove's; the other ranks get *proof_out = NULL.  A failure of
 * a rank-local step (an unsatisfied AIR, say) is returned on EVERY rank.  pipeline_groups: column groups per rank whose exchange
 * runs under the next group's upload + transforms (0 = default: 2 from 2^20 rows on, else 1). */
int s252_cairo_prove_sharded(s252_ctx *ctx, s252_comm *comm, const s252_cairo_trace *trace, size_t blowup, size_t fri_number_of_queries,
                             uint64_t coset_offset, uint8_t grinding_factor, size_t pipeline_groups, uint8_t **proof_out,
                             size_t *proof_len);
/* StarkProof::serialize (proof/stark.rs:161-218) from pieces the caller assembled (the sharded prover gathers them from several
 * GPUs): ood = the frame (2 x cols, row-major; cols = main_cols + aux_cols), hz = H1(z^2), H2(z^2); evs / ev / pas / pa as
 * s252_fri_query returns them ([Q][layers] values, [Q][layers][depth] digests, layer k uses depth - k of them); rows and paths as
 * s252_commit_open returns them (depth digests per query).  *proof_out is malloc'ed; release it with s252_cairo_proof_free. */
int s252_cairo_serialize_proof(size_t trace_rows, const uint8_t *root_main, const uint8_t *root_aux, const uint8_t *root_comp,
                               const s252_fe *ood, size_t cols, const s252_fe *hz, size_t layers, const uint8_t *fri_roots,
                               const s252_fe *last, size_t n_queries, size_t depth, const s252_fe *evs, const s252_fe *ev,
                               const uint8_t *pas, const uint8_t *pa, const s252_fe *comp_rows, const uint8_t *comp_paths,
                               const s252_fe *main_rows, size_t main_cols, const uint8_t *main_paths, const s252_fe *aux_rows,
                               size_t aux_cols, const uint8_t *aux_paths, uint64_t nonce, uint8_t **proof_out, size_t *proof_len);
/* ---- building blocks of a proof sharded over several GPUs ------------------------------------
 * The same kernels as s252_cairo_prove, applied to this rank's columns (LDE) or to this rank's block of
 * LDE rows; lambdaworks_cairo_prover_b200/cairo_distributed.py strings them together with NCCL.
 * "device columns" / "blocks" are column-major buffers in the library's internal element format
 * (s252_commit_device_lde). */
/* The handle's table column-major (n_cols x n_rows, LW; pinned after s252_cairo_trace_pin). */
const s252_fe *s252_cairo_trace_columns(const s252_cairo_trace *t);
/* compute_trace_polys + compute_lde_trace_evaluations (prover.rs:161-185) for n_cols columns given
 * column-major in (pinned) host memory / on the device; no tree. */
int s252_lde_host_columns(s252_ctx *ctx, const s252_fe *cols_lw, size_t n_rows, size_t n_cols, size_t blowup,
                          uint64_t coset_offset, int keep_trace, s252_commit **out);
/* The trace evaluations a handle kept (keep_trace / s252_cairo_round1): device columns [n_cols][n_coeffs], or NULL. */
const void *s252_commit_device_trace(const s252_commit *c);
int s252_lde_device_columns(s252_ctx *ctx, const void *cols, size_t n_rows, size_t n_cols, size_t blowup,
                            uint64_t coset_offset, s252_commit **out);
/* build_auxiliary_trace (air.rs:660-729) on this device: *aux_out = device columns [18][n_rows]
 * (release with s252_device_free).  It reads trace columns 19..29 (pc .. off_op1): `prefetched`, if not
 * NULL, holds those 11 columns on the device -- either a copy of s252_cairo_trace_columns(t) + 19*n_rows
 * (LW elements, prefetched_internal = 0) or device columns in the internal format, e.g. collected from
 * the ranks that own them (s252_commit_device_trace, prefetched_internal = 1); NULL uploads them here. */
int s252_cairo_aux_trace_device(s252_ctx *ctx, const s252_cairo_trace *trace, const s252_fe rap[3], const void *prefetched,
                                int prefetched_internal, void **aux_out);
/* ConstraintEvaluator::evaluate (evaluator.rs:40-262) on LDE rows [row0, row0+rows): main_block / aux_block
 * hold those rows of every column (stride elements apart); *_halo hold the `blowup` rows that follow
 * the block (mod the domain; the frame's next row).  evals_out: device, [rows]. */
int s252_cairo_constraints_rows(s252_ctx *ctx, const s252_cairo_trace *trace, const void *main_block, const void *aux_block,
                                size_t stride, size_t row0, size_t rows, const void *main_halo, const void *aux_halo,
                                size_t halo_stride, const s252_fe rap[3], const s252_fe *boundary_coeffs,
                                const s252_fe *transition_coeffs, size_t blowup, uint64_t coset_offset, void *evals_out);
/* The rest of round 2 (prover.rs:246-283) from the n_rows*blowup constraint evaluations (device):
 * interpolate_offset_fft, even/odd split, LDE of H1/H2, batch_commit. */
int s252_cairo_composition_commit(s252_ctx *ctx, const void *evals, size_t n_rows, size_t blowup, uint64_t coset_offset,
                                  s252_commit **out, uint8_t root[32]);
/* The same without the tree (a rank of a sharded proof hashes only its own block of rows of (H1, H2)). */
int s252_cairo_composition_lde(s252_ctx *ctx, const void *evals, size_t n_rows, size_t blowup, uint64_t coset_offset,
                               s252_commit **out);
/* The DEEP composition polynomial (prover.rs:410-482 as the verifier's formula, verifier.rs:526-557) on
 * LDE rows [row0, row0+rows): tables[t] = block of table t (trace tables first, (H1, H2) last).
 * Argument meaning as in s252_fri_commit_phase_deep.  out: device, [rows]. */
int s252_deep_rows(s252_ctx *ctx, const void *const *tables, const size_t *strides, const size_t *n_cols, size_t n_tables,
                   size_t row0, size_t rows, size_t lde_rows, size_t trace_rows, const s252_fe *z,
                   const uint64_t *transition_offsets, size_t n_offsets, const s252_fe *trace_ood, const s252_fe *h1_z2,
                   const s252_fe *h2_z2, const s252_fe *gamma, const s252_fe *gamma_p, const s252_fe *trace_gammas,
                   uint64_t coset_offset, void *out);
/* Diagnostics: host wall-clock milliseconds per stage of the last s252_cairo_prove on this thread, as JSON. */
const char *s252_cairo_last_prove_stages(void);

#ifdef __cplusplus
}
#endif
#endif
