/*
 * stark252_oracle.c -- CPU ORACLE (test infrastructure, NOT the product).  See the header.
 *
 * Restates, in plain C, what the reference does on the LDE + commitment path.  Each function
 * cites the reference call site (paths relative to /root/reference) or, for behaviour that lives
 * in the un-vendored dependency crates, the published algorithm ("[dep]").
 */
#include "stark252_oracle.h"

#include <stdlib.h>
#include <string.h>
#include <pthread.h>

typedef unsigned __int128 u128;

/* ------------------------------------------------------------------------------------------
 * Field: Stark252PrimeField, Montgomery form with R = 2^256.  [dep] lambdaworks-math
 * field/fields/montgomery_backed_prime_fields.rs, fft_friendly/stark_252_prime_field.rs.
 * Internal element: 4 x u64 little-endian limbs.
 * ------------------------------------------------------------------------------------------ */
typedef struct { uint64_t l[4]; } fe;

/* p = 2^251 + 17*2^192 + 1 */
static const fe P   = {{0x0000000000000001ULL, 0, 0, 0x0800000000000011ULL}};
/* R mod p and R^2 mod p */
static const fe ONE = {{0xffffffffffffffe1ULL, 0xffffffffffffffffULL, 0xffffffffffffffffULL, 0x07fffffffffffdf0ULL}};
static const fe R2  = {{0xfffffd737e000401ULL, 0x00000001330fffffULL, 0xffffffffff6f8000ULL, 0x07ffd4ab5e008810ULL}};
static const fe ZERO = {{0, 0, 0, 0}};
/* [dep] TWO_ADIC_PRIMITVE_ROOT_OF_UNITY (canonical), TWO_ADICITY = 192 */
static const fe W_CANON = {{0x6070024f42f8ef94ULL, 0xad187148e11a6161ULL, 0x3f0464519c8b0fa5ULL, 0x005282db87529cfaULL}};
#define TWO_ADICITY 192

static inline int fe_geq_p(const fe *a) {
    for (int i = 3; i >= 0; --i) {
        if (a->l[i] > P.l[i]) return 1;
        if (a->l[i] < P.l[i]) return 0;
    }
    return 1;
}
static inline void fe_sub_p(fe *a) {
    u128 br = 0;
    for (int i = 0; i < 4; ++i) {
        u128 d = (u128)a->l[i] - P.l[i] - br;
        a->l[i] = (uint64_t)d;
        br = (d >> 64) & 1;
    }
}
static inline fe fe_add(const fe *a, const fe *b) {
    fe r; u128 c = 0;
    for (int i = 0; i < 4; ++i) { c += (u128)a->l[i] + b->l[i]; r.l[i] = (uint64_t)c; c >>= 64; }
    /* a,b < p < 2^252 so no carry out of 256 bits */
    if (fe_geq_p(&r)) fe_sub_p(&r);
    return r;
}
static inline fe fe_sub(const fe *a, const fe *b) {
    fe r; u128 br = 0;
    for (int i = 0; i < 4; ++i) {
        u128 d = (u128)a->l[i] - b->l[i] - br;
        r.l[i] = (uint64_t)d; br = (d >> 64) & 1;
    }
    if (br) { u128 c = 0; for (int i = 0; i < 4; ++i) { c += (u128)r.l[i] + P.l[i]; r.l[i] = (uint64_t)c; c >>= 64; } }
    return r;
}
/* CIOS Montgomery multiplication, mu = -p^-1 mod 2^64 = 2^64-1 (p = 1 mod 2^64). */
static inline fe fe_mul(const fe *a, const fe *b) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; ++i) {
        u128 c = 0;
        for (int j = 0; j < 4; ++j) {
            c += (u128)a->l[j] * b->l[i] + t[j];
            t[j] = (uint64_t)c; c >>= 64;
        }
        c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
        uint64_t m = (uint64_t)(0 - t[0]);   /* t[0] * mu mod 2^64 */
        c = (u128)m * P.l[0] + t[0]; c >>= 64;
        for (int j = 1; j < 4; ++j) {
            c += (u128)m * P.l[j] + t[j];
            t[j - 1] = (uint64_t)c; c >>= 64;
        }
        c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
    }
    fe r = {{t[0], t[1], t[2], t[3]}};
    if (t[4] || fe_geq_p(&r)) fe_sub_p(&r);
    return r;
}
static inline fe fe_to_mont(const fe *canon) { return fe_mul(canon, &R2); }
static inline fe fe_from_mont(const fe *m) { fe one = {{1, 0, 0, 0}}; return fe_mul(m, &one); }
static inline int fe_is_zero(const fe *a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
static fe fe_pow_u64(const fe *a, uint64_t e) {
    fe r = ONE, b = *a;
    while (e) { if (e & 1) r = fe_mul(&r, &b); b = fe_mul(&b, &b); e >>= 1; }
    return r;
}
static fe fe_inv(const fe *a) {
    /* a^(p-2) */
    fe e = P;                 /* p-2: the low limb is 1, so the subtraction borrows */
    { u128 br = 2; for (int i = 0; i < 4; ++i) { u128 d = (u128)e.l[i] - (uint64_t)br; e.l[i] = (uint64_t)d; br = (d >> 64) & 1; } }
    fe r = ONE, b = *a;
    for (int i = 0; i < 4; ++i)
        for (int k = 0; k < 64; ++k) {
            if ((e.l[i] >> k) & 1) r = fe_mul(&r, &b);
            b = fe_mul(&b, &b);
        }
    return r;
}
static inline fe fe_from_u64(uint64_t v) { fe c = {{v, 0, 0, 0}}; return fe_to_mont(&c); }

/* LW <-> internal */
static inline fe lw_in(const fe_lw *a) { fe r = {{a->limbs[3], a->limbs[2], a->limbs[1], a->limbs[0]}}; return r; }
static inline void lw_out(const fe *a, fe_lw *o) { o->limbs[0] = a->l[3]; o->limbs[1] = a->l[2]; o->limbs[2] = a->l[1]; o->limbs[3] = a->l[0]; }

/* [dep] ByteConversion::to_bytes_be: canonical representative, 32 bytes big-endian */
static inline void fe_to_be(const fe *m, uint8_t out[32]) {
    fe c = fe_from_mont(m);
    for (int i = 0; i < 4; ++i)
        for (int k = 0; k < 8; ++k) out[8 * i + k] = (uint8_t)(c.l[3 - i] >> (56 - 8 * k));
}
static inline fe fe_from_be(const uint8_t in[32]) {
    fe c;
    for (int i = 0; i < 4; ++i) {
        uint64_t v = 0;
        for (int k = 0; k < 8; ++k) v = (v << 8) | in[8 * i + k];
        c.l[3 - i] = v;
    }
    return fe_to_mont(&c);   /* Montgomery multiplication by R^2 also reduces values >= p */
}

void o_fe_from_u64(uint64_t v, fe_lw *out) { fe r = fe_from_u64(v); lw_out(&r, out); }
void o_fe_from_bytes_be(const uint8_t in[32], fe_lw *out) { fe r = fe_from_be(in); lw_out(&r, out); }
void o_fe_to_bytes_be(const fe_lw *a, uint8_t out[32]) { fe x = lw_in(a); fe_to_be(&x, out); }
void o_fe_add(const fe_lw *a, const fe_lw *b, fe_lw *out) { fe x = lw_in(a), y = lw_in(b), r = fe_add(&x, &y); lw_out(&r, out); }
void o_fe_sub(const fe_lw *a, const fe_lw *b, fe_lw *out) { fe x = lw_in(a), y = lw_in(b), r = fe_sub(&x, &y); lw_out(&r, out); }
void o_fe_mul(const fe_lw *a, const fe_lw *b, fe_lw *out) { fe x = lw_in(a), y = lw_in(b), r = fe_mul(&x, &y); lw_out(&r, out); }
void o_fe_inv(const fe_lw *a, fe_lw *out) { fe x = lw_in(a), r = fe_inv(&x); lw_out(&r, out); }
void o_fe_pow(const fe_lw *a, uint64_t e, fe_lw *out) { fe x = lw_in(a), r = fe_pow_u64(&x, e); lw_out(&r, out); }

/* [dep] IsFFTField::get_primitive_root_of_unity(order) = W^(2^(TWO_ADICITY-order)); call sites
 * src/starks/domain.rs:30, src/starks/verifier.rs:367. */
static int primitive_root(uint32_t order, fe *out) {
    if (order > TWO_ADICITY) return -1;
    fe w = fe_to_mont(&W_CANON);
    for (uint32_t i = 0; i < TWO_ADICITY - order; ++i) w = fe_mul(&w, &w);
    *out = w;
    return 0;
}
int o_primitive_root(uint32_t order, fe_lw *out) {
    fe w; if (primitive_root(order, &w)) return -1; lw_out(&w, out); return 0;
}
/* [dep] get_powers_of_primitive_root_coset; call sites src/starks/domain.rs:31,39 */
int o_coset_powers(uint32_t order, size_t count, const fe_lw *offset, fe_lw *out) {
    fe w; if (primitive_root(order, &w)) return -1;
    fe cur = lw_in(offset);
    for (size_t i = 0; i < count; ++i) { lw_out(&cur, &out[i]); cur = fe_mul(&cur, &w); }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Keccak-256 (sha3 crate 0.10.6, Cargo.toml:17): Keccak-f[1600], rate 136, pad 0x01 .. 0x80.
 * ------------------------------------------------------------------------------------------ */
static const uint64_t KRC[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
    0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
    0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
    0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
    0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
static const int KROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
static inline uint64_t rol64(uint64_t x, int n) { return n ? (x << n) | (x >> (64 - n)) : x; }
static void keccak_f(uint64_t s[25]) {
    for (int r = 0; r < 24; ++r) {
        uint64_t C[5], D[5], B[25];
        for (int x = 0; x < 5; ++x) C[x] = s[x] ^ s[x + 5] ^ s[x + 10] ^ s[x + 15] ^ s[x + 20];
        for (int x = 0; x < 5; ++x) D[x] = C[(x + 4) % 5] ^ rol64(C[(x + 1) % 5], 1);
        for (int i = 0; i < 25; ++i) s[i] ^= D[i % 5];
        for (int x = 0; x < 5; ++x)
            for (int y = 0; y < 5; ++y)
                B[y + 5 * ((2 * x + 3 * y) % 5)] = rol64(s[x + 5 * y], KROT[x + 5 * y]);
        for (int y = 0; y < 5; ++y)
            for (int x = 0; x < 5; ++x)
                s[x + 5 * y] = B[x + 5 * y] ^ (~B[(x + 1) % 5 + 5 * y] & B[(x + 2) % 5 + 5 * y]);
        s[0] ^= KRC[r];
    }
}
typedef struct { uint64_t s[25]; uint8_t buf[136]; size_t fill; } keccak_ctx;
static void keccak_init(keccak_ctx *c) { memset(c, 0, sizeof *c); }
static void keccak_absorb_block(keccak_ctx *c, const uint8_t *b) {
    for (int i = 0; i < 17; ++i) {
        uint64_t v; memcpy(&v, b + 8 * i, 8);   /* little-endian host */
        c->s[i] ^= v;
    }
    keccak_f(c->s);
}
static void keccak_update(keccak_ctx *c, const uint8_t *d, size_t n) {
    while (n) {
        size_t take = 136 - c->fill; if (take > n) take = n;
        memcpy(c->buf + c->fill, d, take); c->fill += take; d += take; n -= take;
        if (c->fill == 136) { keccak_absorb_block(c, c->buf); c->fill = 0; }
    }
}
static void keccak_final(keccak_ctx *c, uint8_t out[32]) {
    memset(c->buf + c->fill, 0, 136 - c->fill);
    c->buf[c->fill] ^= 0x01; c->buf[135] ^= 0x80;
    keccak_absorb_block(c, c->buf);
    memcpy(out, c->s, 32);
}
void o_keccak256(const uint8_t *data, size_t len, uint8_t out[32]) {
    keccak_ctx c; keccak_init(&c); keccak_update(&c, data, len); keccak_final(&c, out);
}

/* ------------------------------------------------------------------------------------------
 * DefaultTranscript [dep] lambdaworks-crypto fiat_shamir/default_transcript.rs; call sites
 * src/starks/prover.rs:91-94, verifier.rs:37-40.  append = absorb; challenge = finalize,
 * REVERSE the 32 bytes, reset, absorb the reversed bytes, return them.
 * ------------------------------------------------------------------------------------------ */
struct o_transcript { keccak_ctx k; };
o_transcript *o_transcript_new(void) { o_transcript *t = malloc(sizeof *t); keccak_init(&t->k); return t; }
void o_transcript_free(o_transcript *t) { free(t); }
void o_transcript_append(o_transcript *t, const uint8_t *data, size_t len) { keccak_update(&t->k, data, len); }
void o_transcript_challenge(o_transcript *t, uint8_t out[32]) {
    uint8_t h[32];
    keccak_final(&t->k, h);
    for (int i = 0; i < 32; ++i) out[i] = h[31 - i];
    keccak_init(&t->k);
    keccak_update(&t->k, out, 32);
}
/* src/starks/transcript.rs:13-43: clear the top 256-251 = 5 bits, interpret big-endian. */
static fe randomness_to_field(uint8_t r[32]) {
    /* transcript.rs:23-43 with field_bit_size = 252: 256 - 251 = 5 bits to clear -> mask 0x07 */
    r[0] &= 0x07;
    return fe_from_be(r);
}
void o_randomness_to_field(const uint8_t in[32], fe_lw *out) {
    uint8_t r[32]; memcpy(r, in, 32);
    fe v = randomness_to_field(r); lw_out(&v, out);
}
static fe transcript_to_field(o_transcript *t) {
    uint8_t r[32]; o_transcript_challenge(t, r);
    return randomness_to_field(r);
}
void o_transcript_to_field(o_transcript *t, fe_lw *out) { fe r = transcript_to_field(t); lw_out(&r, out); }
/* src/starks/transcript.rs:45-51 */
uint64_t o_transcript_to_usize(o_transcript *t) {
    uint8_t r[32]; o_transcript_challenge(t, r);
    uint64_t v = 0; for (int i = 0; i < 8; ++i) v = (v << 8) | r[i];
    return v;
}

/* ------------------------------------------------------------------------------------------
 * NTT.  [dep] lambdaworks-math fft/cpu: radix-2, results in natural order.  Any correct
 * algorithm yields the same canonical values; this one is iterative Cooley-Tukey with a
 * precomputed twiddle table.
 * ------------------------------------------------------------------------------------------ */
static inline uint32_t ilog2(size_t n) { uint32_t k = 0; while (((size_t)1 << k) < n) ++k; return k; }
static inline size_t next_pow2(size_t n) { size_t r = 1; while (r < n) r <<= 1; return r; }
static inline size_t bitrev(size_t x, uint32_t bits) {
    size_t r = 0; for (uint32_t i = 0; i < bits; ++i) { r = (r << 1) | (x & 1); x >>= 1; } return r;
}
/* tw[i] = w^i for i < n/2 */
static fe *make_twiddles(size_t n, int inverse) {
    uint32_t k = ilog2(n);
    fe w; if (primitive_root(k, &w)) return NULL;
    if (inverse) w = fe_inv(&w);
    size_t h = n / 2 ? n / 2 : 1;
    fe *tw = malloc(h * sizeof(fe));
    fe cur = ONE;
    for (size_t i = 0; i < h; ++i) { tw[i] = cur; cur = fe_mul(&cur, &w); }
    return tw;
}
/* in-place, natural in -> natural out */
static void ntt_inplace(fe *a, size_t n, const fe *tw) {
    uint32_t k = ilog2(n);
    for (size_t i = 0; i < n; ++i) { size_t j = bitrev(i, k); if (i < j) { fe t = a[i]; a[i] = a[j]; a[j] = t; } }
    for (size_t len = 2; len <= n; len <<= 1) {
        size_t half = len / 2, step = n / len;
        for (size_t blk = 0; blk < n; blk += len)
            for (size_t j = 0; j < half; ++j) {
                fe u = a[blk + j];
                fe v = j ? fe_mul(&a[blk + j + half], &tw[j * step]) : a[blk + j + half];
                a[blk + j] = fe_add(&u, &v);
                a[blk + j + half] = fe_sub(&u, &v);
            }
    }
}
static int intt_scaled(fe *a, size_t n) {
    fe *tw = make_twiddles(n, 1); if (!tw) return -1;
    ntt_inplace(a, n, tw); free(tw);
    fe ninv = fe_from_u64((uint64_t)n); ninv = fe_inv(&ninv);
    for (size_t i = 0; i < n; ++i) a[i] = fe_mul(&a[i], &ninv);
    return 0;
}
static inline int is_pow2(size_t n) { return n && !(n & (n - 1)); }

/* Polynomial::interpolate_fft -- call site src/starks/trace.rs:107.  inverse FFT over {g^i},
 * g = root(log2 n), scaled by 1/n.  (Polynomial::new trims trailing zeros; the caller does.) */
int o_interpolate_fft(const fe_lw *evals, size_t n, fe_lw *coeffs) {
    if (!is_pow2(n)) return -1;
    fe *a = malloc(n * sizeof(fe));
    for (size_t i = 0; i < n; ++i) a[i] = lw_in(&evals[i]);
    int rc = intt_scaled(a, n);
    if (!rc) for (size_t i = 0; i < n; ++i) lw_out(&a[i], &coeffs[i]);
    free(a);
    return rc;
}
/* Polynomial::interpolate_offset_fft -- call site src/starks/constraints/evaluation_table.rs:32:
 * interpolate_fft, then scale coefficient i by offset^-i. */
int o_interpolate_offset_fft(const fe_lw *evals, size_t n, const fe_lw *offset, fe_lw *coeffs) {
    if (!is_pow2(n)) return -1;
    fe *a = malloc(n * sizeof(fe));
    for (size_t i = 0; i < n; ++i) a[i] = lw_in(&evals[i]);
    int rc = intt_scaled(a, n);
    if (!rc) {
        fe off = lw_in(offset), oinv = fe_inv(&off), cur = ONE;
        for (size_t i = 0; i < n; ++i) { a[i] = fe_mul(&a[i], &cur); cur = fe_mul(&cur, &oinv); lw_out(&a[i], &coeffs[i]); }
    }
    free(a);
    return rc;
}
static size_t trimmed_len(const fe_lw *c, size_t n) {
    while (n && (c[n - 1].limbs[0] | c[n - 1].limbs[1] | c[n - 1].limbs[2] | c[n - 1].limbs[3]) == 0) --n;
    return n;
}
/* [dep] FFTPoly::evaluate_offset_fft(blowup, domain_size, offset) -- call sites
 * src/starks/prover.rs:117, src/starks/fri/fri_commitment.rs:36:
 *   scaled = p.scale(offset); len = max(coeff_len, domain_size).next_power_of_two() * blowup;
 *   zero-pad to len, forward FFT, natural order: out[i] = p(offset * w_len^i). */
size_t o_evaluate_offset_fft_len(const fe_lw *coeffs, size_t n_coeffs, size_t blowup, size_t domain_size) {
    size_t cl = trimmed_len(coeffs, n_coeffs);
    size_t m = cl > domain_size ? cl : domain_size;
    return next_pow2(m) * blowup;
}
static int evaluate_offset_fft_core(const fe *coeffs, size_t cl, size_t len, const fe *offset, fe *a) {
    fe cur = ONE;
    for (size_t i = 0; i < cl; ++i) { a[i] = fe_mul(&coeffs[i], &cur); cur = fe_mul(&cur, offset); }
    for (size_t i = cl; i < len; ++i) a[i] = ZERO;
    fe *tw = make_twiddles(len, 0); if (!tw) return -1;
    ntt_inplace(a, len, tw); free(tw);
    return 0;
}
int o_evaluate_offset_fft(const fe_lw *coeffs, size_t n_coeffs, size_t blowup, size_t domain_size,
                          const fe_lw *offset, fe_lw *out) {
    size_t cl = trimmed_len(coeffs, n_coeffs);
    size_t len = o_evaluate_offset_fft_len(coeffs, n_coeffs, blowup, domain_size);
    if (!is_pow2(len)) return -1;
    fe *c = malloc((cl ? cl : 1) * sizeof(fe)), *a = malloc(len * sizeof(fe));
    for (size_t i = 0; i < cl; ++i) c[i] = lw_in(&coeffs[i]);
    fe off = lw_in(offset);
    int rc = evaluate_offset_fft_core(c, cl, len, &off, a);
    if (!rc) for (size_t i = 0; i < len; ++i) lw_out(&a[i], &out[i]);
    free(c); free(a);
    return rc;
}
/* evaluate_polynomial_on_lde_domain -- src/starks/prover.rs:106-123 */
int o_evaluate_polynomial_on_lde_domain(const fe_lw *coeffs, size_t n_coeffs, size_t blowup,
                                        size_t domain_size, const fe_lw *offset, fe_lw *out) {
    size_t len = o_evaluate_offset_fft_len(coeffs, n_coeffs, blowup, domain_size);
    size_t want = domain_size * blowup;
    if (want == 0 || len % want) return -1;
    size_t step = len / want;
    if (step == 1) return o_evaluate_offset_fft(coeffs, n_coeffs, blowup, domain_size, offset, out);
    fe_lw *tmp = malloc(len * sizeof(fe_lw));
    int rc = o_evaluate_offset_fft(coeffs, n_coeffs, blowup, domain_size, offset, tmp);
    if (!rc) for (size_t i = 0; i < want; ++i) out[i] = tmp[i * step];
    free(tmp);
    return rc;
}

/* ------------------------------------------------------------------------------------------
 * Merkle trees.  [dep] lambdaworks-crypto merkle_tree/{merkle,proof}.rs and
 * backends/{batch,field_element}.rs with Keccak256; selected in src/starks/config.rs:10-20.
 *   leaf  = Keccak256(row[0].to_bytes_be() || row[1].to_bytes_be() || ...)
 *   node  = Keccak256(left || right); heap layout, root at index 0, leaf i at n-1+i.
 * ------------------------------------------------------------------------------------------ */
static void hash_row(const fe *row, size_t n_cols, size_t stride, uint8_t out[32]) {
    keccak_ctx c; keccak_init(&c);
    uint8_t b[32];
    for (size_t j = 0; j < n_cols; ++j) { fe_to_be(&row[j * stride], b); keccak_update(&c, b, 32); }
    keccak_final(&c, out);
}
static void build_inner(uint8_t *nodes, size_t n_leaves) {
    for (size_t i = n_leaves - 1; i-- > 0;) {
        keccak_ctx c; keccak_init(&c);
        keccak_update(&c, nodes + 32 * (2 * i + 1), 64);   /* children 2i+1, 2i+2 are adjacent */
        keccak_final(&c, nodes + 32 * i);
    }
}
int o_merkle_build(const fe_lw *rows, size_t n_leaves, size_t n_cols, uint8_t *nodes) {
    if (!is_pow2(n_leaves) || n_cols == 0) return -1;
    fe *row = malloc(n_cols * sizeof(fe));
    for (size_t i = 0; i < n_leaves; ++i) {
        for (size_t j = 0; j < n_cols; ++j) row[j] = lw_in(&rows[i * n_cols + j]);
        hash_row(row, n_cols, 1, nodes + 32 * (n_leaves - 1 + i));
    }
    free(row);
    build_inner(nodes, n_leaves);
    return 0;
}
/* MerkleTree::get_proof_by_pos -- call sites src/starks/prover.rs:500,515, fri/mod.rs:105,107 */
int o_merkle_path(const uint8_t *nodes, size_t n_leaves, size_t pos, uint8_t *path) {
    if (pos >= n_leaves) return -1;
    size_t idx = pos + n_leaves - 1, k = 0;
    while (idx != 0) {
        size_t sib = (idx & 1) ? idx + 1 : idx - 1;
        memcpy(path + 32 * k++, nodes + 32 * sib, 32);
        idx = (idx - 1) / 2;
    }
    return 0;
}
/* Proof::verify -- call sites src/starks/verifier.rs:397,417,501,508 */
int o_merkle_verify(const uint8_t root[32], size_t index, const fe_lw *value, size_t n_cols,
                    const uint8_t *path, size_t path_len) {
    fe *row = malloc(n_cols * sizeof(fe));
    for (size_t j = 0; j < n_cols; ++j) row[j] = lw_in(&value[j]);
    uint8_t h[32], buf[64];
    hash_row(row, n_cols, 1, h); free(row);
    for (size_t k = 0; k < path_len; ++k) {
        if ((index & 1) == 0) { memcpy(buf, h, 32); memcpy(buf + 32, path + 32 * k, 32); }
        else { memcpy(buf, path + 32 * k, 32); memcpy(buf + 32, h, 32); }
        o_keccak256(buf, 64, h);
        index >>= 1;
    }
    return memcmp(h, root, 32) == 0;
}

/* ------------------------------------------------------------------------------------------
 * FRI -- src/starks/fri/{mod,fri_commitment,fri_functions}.rs
 * ------------------------------------------------------------------------------------------ */
static size_t fold_poly(const fe *c, size_t n, const fe *beta, fe *out) {
    /* fri_functions.rs:4-27: even + beta * odd */
    size_t m = (n + 1) / 2;
    for (size_t i = 0; i < m; ++i) {
        fe e = c[2 * i];
        if (2 * i + 1 < n) { fe o = fe_mul(&c[2 * i + 1], beta); out[i] = fe_add(&e, &o); }
        else out[i] = e;
    }
    return m;
}
void o_fold_polynomial(const fe_lw *coeffs, size_t n, const fe_lw *beta, fe_lw *out) {
    fe *c = malloc((n ? n : 1) * sizeof(fe)), *o = malloc(((n + 1) / 2 + 1) * sizeof(fe));
    for (size_t i = 0; i < n; ++i) c[i] = lw_in(&coeffs[i]);
    fe b = lw_in(beta);
    size_t m = fold_poly(c, n, &b, o);
    for (size_t i = 0; i < m; ++i) lw_out(&o[i], &out[i]);
    free(c); free(o);
}
static size_t fe_trimmed(const fe *c, size_t n) { while (n && fe_is_zero(&c[n - 1])) --n; return n; }

/* FriLayer::new (fri_commitment.rs:30-47): evaluate_offset_fft(1, Some(domain_size), offset),
 * then FriMerkleTree::build (single-element leaves). */
static int fri_layer_new(const fe *poly, size_t n, const fe *offset, size_t domain_size,
                         fe_lw *evals_out, uint8_t *nodes_out, uint8_t root[32]) {
    size_t cl = fe_trimmed(poly, n);
    size_t len = next_pow2(cl > domain_size ? cl : domain_size);
    if (len != domain_size) return -2;   /* a degree >= domain_size polynomial: not on the path */
    fe *a = malloc(len * sizeof(fe));
    int rc = evaluate_offset_fft_core(poly, cl, len, offset, a);
    if (rc) { free(a); return rc; }
    uint8_t *nodes = nodes_out ? nodes_out : malloc(32 * (2 * len - 1));
    for (size_t i = 0; i < len; ++i) hash_row(&a[i], 1, 1, nodes + 32 * (len - 1 + i));
    build_inner(nodes, len);
    memcpy(root, nodes, 32);
    if (evals_out) for (size_t i = 0; i < len; ++i) lw_out(&a[i], &evals_out[i]);
    if (!nodes_out) free(nodes);
    free(a);
    return 0;
}
/* fri_commit_phase (fri/mod.rs:20-72) */
int o_fri_commit_phase(size_t number_layers, const fe_lw *p0, size_t n_coeffs, o_transcript *t,
                       const fe_lw *coset_offset, size_t domain_size, fe_lw *last_value,
                       fe_lw **layer_evals, uint8_t **layer_nodes, uint8_t *roots) {
    fe *cur = malloc((n_coeffs ? n_coeffs : 1) * sizeof(fe)), *nxt = malloc((n_coeffs / 2 + 2) * sizeof(fe));
    for (size_t i = 0; i < n_coeffs; ++i) cur[i] = lw_in(&p0[i]);
    size_t n = n_coeffs;
    fe off = lw_in(coset_offset);
    int rc = 0;
    if (number_layers > 0) {
        rc = fri_layer_new(cur, n, &off, domain_size, layer_evals ? layer_evals[0] : NULL,
                           layer_nodes ? layer_nodes[0] : NULL, roots);
        if (rc) goto done;
        o_transcript_append(t, roots, 32);
    }
    for (size_t k = 1; k < number_layers; ++k) {
        fe zeta = transcript_to_field(t);
        off = fe_mul(&off, &off);
        domain_size /= 2;
        n = fold_poly(cur, n, &zeta, nxt);
        { fe *s = cur; cur = nxt; nxt = s; }
        rc = fri_layer_new(cur, n, &off, domain_size, layer_evals ? layer_evals[k] : NULL,
                           layer_nodes ? layer_nodes[k] : NULL, roots + 32 * k);
        if (rc) goto done;
        o_transcript_append(t, roots + 32 * k, 32);
    }
    {
        fe zeta = transcript_to_field(t);
        n = fold_poly(cur, n, &zeta, nxt);
        /* Polynomial::new trims trailing zeros; coefficients().get(0).unwrap_or(zero) */
        size_t cl = fe_trimmed(nxt, n);
        fe lv = cl ? nxt[0] : ZERO;
        lw_out(&lv, last_value);
        uint8_t b[32]; fe_to_be(&lv, b);
        o_transcript_append(t, b, 32);
    }
done:
    free(cur); free(nxt);
    return rc;
}

/* ------------------------------------------------------------------------------------------
 * Grinding -- src/starks/grinding.rs:17-48
 * ------------------------------------------------------------------------------------------ */
uint8_t o_grinding_zeros(const uint8_t challenge[32], uint64_t nonce) {
    uint8_t data[40], d[32];
    memcpy(data, challenge, 32);
    for (int i = 0; i < 8; ++i) data[32 + i] = (uint8_t)(nonce >> (8 * i));   /* to_le_bytes */
    o_keccak256(data, 40, d);
    uint64_t head = 0; for (int i = 0; i < 8; ++i) head = (head << 8) | d[i];  /* from_be_bytes */
    return head ? (uint8_t)__builtin_ctzll(head) : 64;
}
int o_generate_nonce_with_grinding(const uint8_t challenge[32], uint8_t grinding_factor,
                                   uint64_t limit, uint64_t *nonce) {
    for (uint64_t n = 0; n < limit; ++n)
        if (o_grinding_zeros(challenge, n) >= grinding_factor) { *nonce = n; return 1; }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * interpolate_and_commit -- src/starks/prover.rs:126-159
 *   trace.compute_trace_polys()            trace.rs:104-110 (sequential over columns)
 *   compute_lde_trace_evaluations          prover.rs:161-185 (rayon over columns)
 *   TraceTable::new_from_cols(..).rows()   trace.rs:45-59, 69-76
 *   batch_commit                           prover.rs:96-104
 * ------------------------------------------------------------------------------------------ */
int o_commit_columns(const fe_lw *cols, size_t n_rows, size_t n_cols, uint8_t *nodes_out, uint8_t root[32]) {
    if (!is_pow2(n_rows) || n_cols == 0) return -1;
    uint8_t *nodes = nodes_out ? nodes_out : malloc(32 * (2 * n_rows - 1));
    fe *row = malloc(n_cols * sizeof(fe));
    for (size_t i = 0; i < n_rows; ++i) {
        for (size_t j = 0; j < n_cols; ++j) row[j] = lw_in(&cols[j * n_rows + i]);
        hash_row(row, n_cols, 1, nodes + 32 * (n_rows - 1 + i));
    }
    free(row);
    build_inner(nodes, n_rows);
    memcpy(root, nodes, 32);
    if (!nodes_out) free(nodes);
    return 0;
}
/* threads over columns: the granularity the reference has under `parallel` (prover.rs:169-172) */
typedef struct {
    const fe *coeffs; fe *lde; const fe *hp; const fe *twf;
    size_t n_rows, m, n_cols, next;
    pthread_mutex_t mu;
} lde_job;
static void *lde_worker(void *arg) {
    lde_job *jb = arg;
    for (;;) {
        pthread_mutex_lock(&jb->mu);
        size_t j = jb->next++;
        pthread_mutex_unlock(&jb->mu);
        if (j >= jb->n_cols) break;
        const fe *c = jb->coeffs + j * jb->n_rows;
        fe *a = jb->lde + j * jb->m;
        for (size_t i = 0; i < jb->n_rows; ++i) a[i] = fe_mul(&c[i], &jb->hp[i]);
        for (size_t i = jb->n_rows; i < jb->m; ++i) a[i] = ZERO;
        ntt_inplace(a, jb->m, jb->twf);
    }
    return NULL;
}
int o_interpolate_and_commit(const fe_lw *trace, size_t n_rows, size_t n_cols, size_t blowup,
                             uint64_t coset_offset, int threads, fe_lw *coeffs_out, fe_lw *lde_out,
                             uint8_t *nodes_out, uint8_t root[32]) {
    if (!is_pow2(n_rows) || !is_pow2(blowup) || n_cols == 0) return -1;
    size_t m = n_rows * blowup;
    fe *coeffs = malloc(n_cols * n_rows * sizeof(fe));
    fe *lde = malloc(n_cols * m * sizeof(fe));
    if (!coeffs || !lde) { free(coeffs); free(lde); return -3; }
    /* compute_trace_polys: cols() gather + interpolate_fft per column, sequential */
    fe *twi = make_twiddles(n_rows, 1);
    fe ninv = fe_from_u64((uint64_t)n_rows); ninv = fe_inv(&ninv);
    for (size_t j = 0; j < n_cols; ++j) {
        fe *c = coeffs + j * n_rows;
        for (size_t i = 0; i < n_rows; ++i) c[i] = lw_in(&trace[i * n_cols + j]);
        ntt_inplace(c, n_rows, twi);
        for (size_t i = 0; i < n_rows; ++i) c[i] = fe_mul(&c[i], &ninv);
    }
    free(twi);
    /* compute_lde_trace_evaluations: evaluate_offset_fft(blowup, Some(n_rows), h) per column */
    fe h = fe_from_u64(coset_offset);
    fe *twf = make_twiddles(m, 0);
    fe *hp = malloc(n_rows * sizeof(fe));
    { fe cur = ONE; for (size_t i = 0; i < n_rows; ++i) { hp[i] = cur; cur = fe_mul(&cur, &h); } }
    {
        lde_job job;
        job.coeffs = coeffs; job.lde = lde; job.hp = hp; job.twf = twf;
        job.n_rows = n_rows; job.m = m; job.n_cols = n_cols; job.next = 0;
        pthread_mutex_init(&job.mu, NULL);
        int nth = threads > 0 ? threads : 1;
        if (nth > 256) nth = 256;
        pthread_t tid[256];
        for (int i = 1; i < nth; ++i) pthread_create(&tid[i], NULL, lde_worker, &job);
        lde_worker(&job);
        for (int i = 1; i < nth; ++i) pthread_join(tid[i], NULL);
        pthread_mutex_destroy(&job.mu);
    }
    free(twf); free(hp);
    /* new_from_cols + rows + batch_commit: sequential leaf hashing + tree */
    uint8_t *nodes = nodes_out ? nodes_out : malloc(32 * (2 * m - 1));
    fe *row = malloc(n_cols * sizeof(fe));
    for (size_t i = 0; i < m; ++i) {
        for (size_t j = 0; j < n_cols; ++j) row[j] = lde[j * m + i];
        hash_row(row, n_cols, 1, nodes + 32 * (m - 1 + i));
    }
    free(row);
    build_inner(nodes, m);
    memcpy(root, nodes, 32);
    if (coeffs_out) for (size_t i = 0; i < n_cols * n_rows; ++i) lw_out(&coeffs[i], &coeffs_out[i]);
    if (lde_out) for (size_t i = 0; i < n_cols * m; ++i) lw_out(&lde[i], &lde_out[i]);
    if (!nodes_out) free(nodes);
    free(coeffs); free(lde);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Round 3 / round 4 helpers (first "next" row)
 * ------------------------------------------------------------------------------------------ */
/* Polynomial::evaluate (Horner) -- call sites src/starks/prover.rs:296-300, frame.rs:76-80 */
void o_poly_evaluate(const fe_lw *coeffs, size_t n, const fe_lw *x, fe_lw *out) {
    fe acc = ZERO, xx = lw_in(x);
    for (size_t i = n; i-- > 0;) { fe c = lw_in(&coeffs[i]); acc = fe_mul(&acc, &xx); acc = fe_add(&acc, &c); }
    lw_out(&acc, out);
}
/* (p - p(b)) / (X - b) in place on n coefficients -> n-1 coefficients ([dep] ruffini_division_inplace) */
static void ruffini(fe *c, size_t n, const fe *b) {
    if (n == 0) return;
    /* synthetic division from the top: q[i-1] = c[i] + b*q[i]; the remainder c[0] + b*q[0] is dropped */
    fe carry = ZERO;
    for (size_t i = n - 1; i > 0; --i) {
        fe t = fe_mul(&carry, b);
        carry = fe_add(&c[i], &t);
        c[i] = carry;
    }
    for (size_t i = 0; i + 1 < n; ++i) c[i] = c[i + 1];
    c[n - 1] = ZERO;
}
/* compute_deep_composition_poly -- src/starks/prover.rs:410-482.
 * trace_polys: n_cols x n coefficients (column-major); h1, h2: n coefficients each (zero padded);
 * ood: n_offsets x n_cols (frame rows); gammas: n_cols x n_offsets (index i*n_offsets + k);
 * out: n coefficients of p0. */
void o_deep_composition_poly(const fe_lw *trace_polys, size_t n_cols, size_t n, const fe_lw *h1, const fe_lw *h2,
                             const fe_lw *z, const uint64_t *offsets, size_t n_offsets, const fe_lw *ood,
                             const fe_lw *h1_z2, const fe_lw *h2_z2, const fe_lw *gamma, const fe_lw *gamma_p,
                             const fe_lw *gammas, fe_lw *out) {
    fe *acc = calloc(n, sizeof(fe)), *tmp = malloc(n * sizeof(fe));
    fe zz = lw_in(z), z2 = fe_mul(&zz, &zz);
    uint32_t order = ilog2(n);
    fe g = ONE; primitive_root(order, &g);
    /* gamma (H1 - H1(z^2)) / (X - z^2) + gamma' (H2 - H2(z^2)) / (X - z^2) */
    for (int which = 0; which < 2; ++which) {
        const fe_lw *hh = which ? h2 : h1;
        fe v = lw_in(which ? h2_z2 : h1_z2), gm = lw_in(which ? gamma_p : gamma);
        for (size_t i = 0; i < n; ++i) tmp[i] = lw_in(&hh[i]);
        tmp[0] = fe_sub(&tmp[0], &v);
        ruffini(tmp, n, &z2);
        for (size_t i = 0; i < n; ++i) { fe t = fe_mul(&tmp[i], &gm); acc[i] = fe_add(&acc[i], &t); }
    }
    /* sum_jk gamma_jk (t_j - t_j(z g^k)) / (X - z g^k) */
    for (size_t j = 0; j < n_cols; ++j)
        for (size_t k = 0; k < n_offsets; ++k) {
            fe gk = fe_pow_u64(&g, offsets[k]), zs = fe_mul(&zz, &gk);
            fe v = lw_in(&ood[k * n_cols + j]), gm = lw_in(&gammas[j * n_offsets + k]);
            for (size_t i = 0; i < n; ++i) tmp[i] = lw_in(&trace_polys[j * n + i]);
            tmp[0] = fe_sub(&tmp[0], &v);
            ruffini(tmp, n, &zs);
            for (size_t i = 0; i < n; ++i) { fe t = fe_mul(&tmp[i], &gm); acc[i] = fe_add(&acc[i], &t); }
        }
    for (size_t i = 0; i < n; ++i) lw_out(&acc[i], &out[i]);
    free(acc); free(tmp);
}

/* Cairo AIR pieces (aux trace, constraint evaluation): same translation unit */
#include "cairo_oracle.inc.c"
