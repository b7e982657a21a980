/*
 * stark252_oracle.h -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * A plain-C restatement of the lambdaworks Cairo prover's LDE + commitment hot path
 * (SURVEY.md section 8).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product path
 * (lambdaworks_cairo_prover_b200/) never does.
 *
 * PARITY STATUS: pinned.  The arithmetic lives in two un-vendored crates
 * (lambdaworks-math / lambdaworks-crypto @ a17b951, sha3 0.10.6; Cargo.toml:11-17 of the
 * reference), so their published algorithms are restated here and anchored on the reference's
 * own known-answer tests and golden proof files (tests/test_oracle_golden.py):
 *   - grinding KAT        src/starks/grinding.rs:56-64      (nonce 33)
 *   - field KAT           src/cairo/air.rs:1412-1451
 *   - fold KAT            src/starks/fri/fri_functions.rs:38-63 (shape; over Stark252 here)
 *   - golden proofs       benches/proofs/fibonacci_{500,1000,70000}.proof: every Merkle opening
 *                         (both tree kinds), every FRI fold, the Fiat-Shamir chain, the stored
 *                         nonce (minimal) and the query indices are reproduced.
 *
 * Element interchange format ("LW"): exactly the in-memory FieldElement<Stark252PrimeField> of the
 * reference's dependency -- 4 x u64, limbs[0] MOST significant, value in Montgomery form
 * (R = 2^256), fully reduced.  All fe_lw* pointers below use it.
 */
#ifndef STARK252_ORACLE_H
#define STARK252_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint64_t limbs[4]; } fe_lw;   /* limbs[0] most significant, Montgomery form */

/* ---- field (lambdaworks-math: montgomery_backed_prime_fields, stark_252_prime_field) ---- */
void o_fe_from_u64(uint64_t v, fe_lw *out);
void o_fe_from_bytes_be(const uint8_t in[32], fe_lw *out);  /* canonical BE -> LW (reduces mod p) */
void o_fe_to_bytes_be(const fe_lw *a, uint8_t out[32]);     /* LW -> canonical BE */
void o_fe_add(const fe_lw *a, const fe_lw *b, fe_lw *out);
void o_fe_sub(const fe_lw *a, const fe_lw *b, fe_lw *out);
void o_fe_mul(const fe_lw *a, const fe_lw *b, fe_lw *out);
void o_fe_inv(const fe_lw *a, fe_lw *out);
void o_fe_pow(const fe_lw *a, uint64_t e, fe_lw *out);
/* F::get_primitive_root_of_unity(order): W^(2^(192-order)) */
int  o_primitive_root(uint32_t order, fe_lw *out);
/* get_powers_of_primitive_root_coset(order, count, offset): offset * w^i, natural order */
int  o_coset_powers(uint32_t order, size_t count, const fe_lw *offset, fe_lw *out);

/* ---- Keccak-256 (sha3 crate; 0x01 padding) ---- */
void o_keccak256(const uint8_t *data, size_t len, uint8_t out[32]);

/* ---- DefaultTranscript (lambdaworks-crypto fiat_shamir::default_transcript) ---- */
typedef struct o_transcript o_transcript;
o_transcript *o_transcript_new(void);
void o_transcript_free(o_transcript *t);
void o_transcript_append(o_transcript *t, const uint8_t *data, size_t len);
void o_transcript_challenge(o_transcript *t, uint8_t out[32]);
/* src/starks/transcript.rs:13-51 */
void o_randomness_to_field(const uint8_t in[32], fe_lw *out);
void o_transcript_to_field(o_transcript *t, fe_lw *out);
uint64_t o_transcript_to_usize(o_transcript *t);

/* ---- FFTPoly (lambdaworks-math fft::polynomial) ---- */
/* Polynomial::interpolate_fft: n evals (n = 2^k) on {g^i} -> n coefficients (untrimmed). */
int o_interpolate_fft(const fe_lw *evals, size_t n, fe_lw *coeffs);
/* Polynomial::interpolate_offset_fft */
int o_interpolate_offset_fft(const fe_lw *evals, size_t n, const fe_lw *offset, fe_lw *coeffs);
/* Length of the output of evaluate_offset_fft for a polynomial with n_coeffs coefficients
 * (trailing zeros are trimmed first, as Polynomial::new does). domain_size = 0 means None. */
size_t o_evaluate_offset_fft_len(const fe_lw *coeffs, size_t n_coeffs, size_t blowup, size_t domain_size);
int o_evaluate_offset_fft(const fe_lw *coeffs, size_t n_coeffs, size_t blowup, size_t domain_size,
                          const fe_lw *offset, fe_lw *out);
/* src/starks/prover.rs:106-123 (evaluate_polynomial_on_lde_domain incl. the step rule);
 * out has domain_size*blowup elements. */
int o_evaluate_polynomial_on_lde_domain(const fe_lw *coeffs, size_t n_coeffs, size_t blowup,
                                        size_t domain_size, const fe_lw *offset, fe_lw *out);

/* ---- MerkleTree<B> (lambdaworks-crypto merkle_tree) ---- */
/* nodes: (2*n_leaves-1) x 32 bytes, heap layout, root at 0, leaf i at n_leaves-1+i.
 * rows: row-major n_leaves x n_cols LW elements. n_cols = 1 and batched = 0 gives the FRI tree
 * (Keccak256Tree); batched = 1 gives BatchKeccak256Tree (same bytes for n_cols = 1). */
int o_merkle_build(const fe_lw *rows, size_t n_leaves, size_t n_cols, uint8_t *nodes);
/* get_proof_by_pos: path has log2(n_leaves) x 32 bytes, leaf -> root. */
int o_merkle_path(const uint8_t *nodes, size_t n_leaves, size_t pos, uint8_t *path);
/* Proof::verify */
int o_merkle_verify(const uint8_t root[32], size_t index, const fe_lw *value, size_t n_cols,
                    const uint8_t *path, size_t path_len);

/* ---- FRI (src/starks/fri) ---- */
/* fold_polynomial (fri_functions.rs:4-27): n coeffs -> ceil(n/2) coeffs (untrimmed). */
void o_fold_polynomial(const fe_lw *coeffs, size_t n, const fe_lw *beta, fe_lw *out);
/* fri_commit_phase (fri/mod.rs:20-72). p0: n_coeffs coefficients. Layers k = 0..number_layers-1 of
 * size domain_size >> k.  layer_evals[k] / layer_nodes[k] must be preallocated
 * ((domain_size>>k) elements, 2*(domain_size>>k)-1 digests); either array may be NULL. */
int o_fri_commit_phase(size_t number_layers, const fe_lw *p0, size_t n_coeffs, o_transcript *t,
                       const fe_lw *coset_offset, size_t domain_size, fe_lw *last_value,
                       fe_lw **layer_evals, uint8_t **layer_nodes, uint8_t *roots /* number_layers x 32 */);

/* ---- grinding (src/starks/grinding.rs) ---- */
uint8_t o_grinding_zeros(const uint8_t challenge[32], uint64_t nonce);
/* returns 1 and sets *nonce when found below `limit`, else 0 */
int o_generate_nonce_with_grinding(const uint8_t challenge[32], uint8_t grinding_factor,
                                   uint64_t limit, uint64_t *nonce);

/* ---- interpolate_and_commit (src/starks/prover.rs:126-159) ----
 * trace: row-major n_rows x n_cols.  Outputs (any may be NULL): coeffs column-major
 * n_cols x n_rows; lde column-major n_cols x (n_rows*blowup); nodes of the batched tree over the
 * LDE rows; root.  threads = number of OpenMP threads over columns for the LDE step (the
 * reference's `parallel` granularity, prover.rs:169-172); interpolation and tree build are
 * sequential as in the reference. */
int o_interpolate_and_commit(const fe_lw *trace, size_t n_rows, size_t n_cols, size_t blowup,
                             uint64_t coset_offset, int threads, fe_lw *coeffs, fe_lw *lde,
                             uint8_t *nodes, uint8_t root[32]);

/* Commit to pre-evaluated columns (column-major n_cols x n_rows): batch_commit(prover.rs:96-104)
 * after new_from_cols + rows(). */
int o_commit_columns(const fe_lw *cols, size_t n_rows, size_t n_cols, uint8_t *nodes, uint8_t root[32]);

/* ---- round 3 / round 4 helpers ---- */
/* Polynomial::evaluate (Horner) */
void o_poly_evaluate(const fe_lw *coeffs, size_t n, const fe_lw *x, fe_lw *out);
/* compute_deep_composition_poly (src/starks/prover.rs:410-482), coefficient form with Ruffini divisions */
void o_deep_composition_poly(const fe_lw *trace_polys, size_t n_cols, size_t n, const fe_lw *h1, const fe_lw *h2,
                             const fe_lw *z, const uint64_t *offsets, size_t n_offsets, const fe_lw *ood,
                             const fe_lw *h1_z2, const fe_lw *h2_z2, const fe_lw *gamma, const fe_lw *gamma_p,
                             const fe_lw *gammas, fe_lw *out);

/* ---- Cairo AIR (cairo_oracle.inc.c) ---- */
/* build_auxiliary_trace (src/cairo/air.rs:660-729): main row-major n x n_cols -> aux row-major n x 18;
 * public memory in address order; rap = (alpha_memory, z_memory, z_range_check). */
int o_cairo_build_aux_trace(const fe_lw *main, size_t n, size_t n_cols, const uint64_t *pub_addrs, const fe_lw *pub_vals,
                            size_t n_pub, const fe_lw *rap, fe_lw *aux_out);
/* CairoAIR::compute_transition (src/cairo/air.rs:743-767) on one frame (rows cur, nxt of n_cols values) */
void o_cairo_compute_transition(const fe_lw *cur, const fe_lw *nxt, size_t n_cols, const fe_lw *rap, int has_rc, fe_lw *out);
/* ConstraintEvaluator::evaluate for CairoAIR (src/starks/constraints/evaluator.rs:40-262): lde column-major
 * n_cols x (n*blowup); boundary constraints (col, step, value) in CairoAIR::boundary_constraints order;
 * coefficient arrays hold (alpha_k, beta_k) pairs. */
int o_cairo_constraint_evaluations(const fe_lw *lde, size_t n_cols, size_t n, size_t blowup, uint64_t coset_offset, int has_rc,
                                   const fe_lw *rap, size_t n_boundary, const uint64_t *bcols, const uint64_t *bsteps,
                                   const fe_lw *bvalues, const fe_lw *boundary_coeffs, const fe_lw *transition_coeffs,
                                   int threads, fe_lw *out);

#ifdef __cplusplus
}
#endif
#endif
