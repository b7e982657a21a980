// cairo.cuh -- device kernels of the Cairo stages that sit between the LDE + commitment calls of
// `prove::<CairoAIR>`:
//
//   build_auxiliary_trace          src/cairo/air.rs:660-729    stable radix sort by address, permutation columns
//   ConstraintEvaluator::evaluate  src/starks/constraints/evaluator.rs:40-262 with
//   CairoAIR::compute_transition   src/cairo/air.rs:743-767, 869-1160 and boundary_constraints :777-849
//
// No field inversion runs on the device.  The reference batch-inverts (z - a' - alpha v') before the
// running product; here the running product of numerators and the SUFFIX product of denominators are
// two multiplicative scans and  perm[i] = numprefix[i] * densuffix[i+1] / dentotal  needs one host
// inversion.  The boundary zerofiers 1/(x - g^s) for all boundary steps s come from ONE table
// T[i] = 1/(x_i - 1) over the LDE coset by index rotation: x_i - g^s = g^s (x_{i - blowup*s} - 1).
#pragma once
#include "fe.cuh"

namespace s252 {

// ------------------------------------------------------------------------------------------------
// multiplicative scan (inclusive, in place).  reverse: logical element i lives at data[n-1-i], i.e.
// the result is the suffix product.
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__global__ void __launch_bounds__(SCAN_THREADS) scan_mul_tiles(fe* __restrict__ data, unsigned long long n, fe* __restrict__ totals,
                                                                int reverse) {
    __shared__ __align__(16) fe sh[SCAN_THREADS];
    const unsigned t = threadIdx.x;
    const unsigned long long base = (unsigned long long)blockIdx.x * SCAN_TILE + (unsigned long long)t * SCAN_ITEMS;
    fe v[SCAN_ITEMS];
    fe run = fe_one();
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const unsigned long long idx = base + k;
        if (idx < n) run = fe_mul_full(run, ld_fe(data + (reverse ? n - 1 - idx : idx)));
        v[k] = run;
    }
    st_fe(sh + t, run);
    __syncthreads();
    for (unsigned off = 1; off < SCAN_THREADS; off <<= 1) {
        fe y = fe_one();
        const bool has = t >= off;
        if (has) y = ld_fe(sh + t - off);
        __syncthreads();
        if (has) st_fe(sh + t, fe_mul_full(ld_fe(sh + t), y));
        __syncthreads();
    }
    const fe excl = t ? ld_fe(sh + t - 1) : fe_one();
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const unsigned long long idx = base + k;
        if (idx < n) st_fe(data + (reverse ? n - 1 - idx : idx), t ? fe_mul_full(v[k], excl) : v[k]);
    }
    if (t == SCAN_THREADS - 1 && totals) st_fe(totals + blockIdx.x, ld_fe(sh + t));
}
// data[i] *= prefix[tile(i) - 1] for every tile after the first (prefix = inclusive scan of the tile totals)
__global__ void __launch_bounds__(256) scan_mul_apply(fe* __restrict__ data, unsigned long long n, const fe* __restrict__ prefix,
                                                      int reverse) {
    const unsigned long long idx = (unsigned long long)blockIdx.x * 256 + threadIdx.x + SCAN_TILE;
    if (idx >= n) return;
    const unsigned long long tile = idx / SCAN_TILE;
    fe* p = data + (reverse ? n - 1 - idx : idx);
    st_fe(p, fe_mul_full(ld_fe(p), ld_fe(prefix + tile - 1)));
}
// out[i] = pre[i-1] * suf[i+1] * inv_total : element-wise inverses from inclusive prefix / suffix products
__global__ void __launch_bounds__(256) batch_inverse_finish(fe* __restrict__ out, const fe* __restrict__ pre, const fe* __restrict__ suf,
                                                            unsigned long long n, fe inv_total) {
    const unsigned long long i = (unsigned long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    fe r = inv_total;
    if (i) r = fe_mul_full(r, ld_fe(pre + i - 1));
    if (i + 1 < n) r = fe_mul_full(r, ld_fe(suf + i + 1));
    st_fe(out + i, r);
}
// out[i] = in[i] - c  (twice: the same values feed the prefix and the suffix scan)
__global__ void __launch_bounds__(256) sub_const2(const fe* __restrict__ in, fe* __restrict__ a, fe* __restrict__ b, unsigned long long n, fe c) {
    const unsigned long long i = (unsigned long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const fe v = fe_sub_full(ld_fe(in + i), c);
    st_fe(a + i, v);
    st_fe(b + i, v);
}

// ------------------------------------------------------------------------------------------------
// stable LSD radix sort, 8 bits per pass (sort_columns_by_memory_address uses Rust's stable sort_by,
// air.rs:529-533).  A warp owns RS_ITEMS consecutive elements: pass 1 counts its digits, a scan over
// the digit-major (digit, warp) histogram gives every warp its output base per digit, pass 2 re-reads
// the elements in order and ranks equal digits inside each 32-element chunk with __match_any_sync, so
// equal keys keep their input order.
constexpr unsigned RS_ITEMS = 2048;
constexpr unsigned RS_WARPS = 4;

template <typename K>
__global__ void __launch_bounds__(RS_WARPS * 32) rs_histogram(const K* __restrict__ keys, unsigned n, unsigned shift,
                                                              unsigned* __restrict__ hist, unsigned units) {
    __shared__ unsigned cnt[RS_WARPS][256];
    const unsigned w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned unit = blockIdx.x * RS_WARPS + w;
    if (unit >= units) return;                       // whole warps leave together
    for (unsigned d = lane; d < 256; d += 32) cnt[w][d] = 0;
    __syncwarp();
    const unsigned lo = unit * RS_ITEMS;
    for (unsigned c = 0; c < RS_ITEMS && lo + c < n; c += 32) {
        const unsigned i = lo + c + lane;
        const bool valid = i < n;
        const unsigned d = valid ? ((unsigned)(keys[i] >> shift) & 255u) : 256u + lane;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        if (valid && lane == (unsigned)__ffs(peers) - 1) cnt[w][d] += __popc(peers);      // one update per distinct digit
        __syncwarp();
    }
    for (unsigned d = lane; d < 256; d += 32) hist[d * units + unit] = cnt[w][d];
}
// block d: exclusive scan of row d of the digit-major histogram (in place) and the row total
__global__ void __launch_bounds__(1024) rs_scan_rows(unsigned* __restrict__ hist, unsigned units, unsigned* __restrict__ digit_totals) {
    __shared__ unsigned wsum[32];
    const unsigned t = threadIdx.x, lane = t & 31, wid = t >> 5;
    unsigned* row = hist + blockIdx.x * units;
    unsigned carry = 0;
    for (unsigned base = 0; base < units; base += 1024) {
        const unsigned x = base + t < units ? row[base + t] : 0;
        unsigned incl = x;
        for (unsigned o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
        if (lane == 31) wsum[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            unsigned v = wsum[lane];
            for (unsigned o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += y; }
            wsum[lane] = v;
        }
        __syncthreads();
        const unsigned prefix = wid ? wsum[wid - 1] : 0;
        if (base + t < units) row[base + t] = carry + prefix + incl - x;
        carry += wsum[31];
        __syncthreads();
    }
    if (t == 0) digit_totals[blockIdx.x] = carry;
}
template <typename K, bool HAS_VALS>
__global__ void __launch_bounds__(RS_WARPS * 32) rs_scatter(const K* __restrict__ keys, const unsigned* __restrict__ vals, unsigned n,
                                                            unsigned shift, const unsigned* __restrict__ hist, unsigned units,
                                                            const unsigned* __restrict__ digit_totals, K* __restrict__ keys_out,
                                                            unsigned* __restrict__ vals_out) {
    __shared__ unsigned base[RS_WARPS][256];
    __shared__ unsigned dbase[256];
    const unsigned w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned unit = blockIdx.x * RS_WARPS + w;
    if (w == 0) {                                    // exclusive scan of the 256 digit totals: 8 digits per lane
        unsigned v[8], s = 0;
        for (int k = 0; k < 8; ++k) { v[k] = digit_totals[lane * 8 + k]; s += v[k]; }
        unsigned incl = s;
        for (unsigned o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
        unsigned run = incl - s;
        for (int k = 0; k < 8; ++k) { dbase[lane * 8 + k] = run; run += v[k]; }
    }
    __syncthreads();
    if (unit >= units) return;
    for (unsigned d = lane; d < 256; d += 32) base[w][d] = dbase[d] + hist[d * units + unit];
    __syncwarp();
    const unsigned lo = unit * RS_ITEMS;
    for (unsigned c = 0; c < RS_ITEMS && lo + c < n; c += 32) {
        const unsigned i = lo + c + lane;
        const bool valid = i < n;
        const K key = valid ? keys[i] : (K)0;
        const unsigned d = (unsigned)(key >> shift) & 255u;
        const unsigned peers = __match_any_sync(0xffffffffu, valid ? d : 256u + lane);
        const unsigned rank = __popc(peers & ((1u << lane) - 1u));
        const unsigned pos = valid ? base[w][d] + rank : 0;
        __syncwarp();
        if (valid && rank + 1 == (unsigned)__popc(peers)) base[w][d] += rank + 1;     // the last peer advances the digit's cursor
        __syncwarp();
        if (valid) {
            keys_out[pos] = key;
            if (HAS_VALS) vals_out[pos] = vals[i];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// auxiliary trace
constexpr unsigned CAIRO_PC = 19, CAIRO_INST = 23, CAIRO_OFF_DST = 27, CAIRO_AUX_COLS = 18;

struct CairoAux {
    const fe* main;                   // column-major main trace (internal format), starting at trace column `col0`
    unsigned col0;                    // 0: the whole table; 19: only the 11 columns pc .. off_op1 the builder reads
    unsigned long long n;             // rows
    const unsigned long long* pub_addr;   // public memory addresses (address order)
    const fe* pub_addr_fe;            // the same as field elements
    const fe* pub_val;
    unsigned n_pub;
    fe alpha, z, zrc;
    fe* aux;                          // column-major [18][n]
};
__device__ __forceinline__ unsigned long long fe_low64(const fe& mont) {
    const fe c = fe_from_mont(mont);
    return (unsigned long long)c.l[0] | ((unsigned long long)c.l[1] << 32);
}
// long-format sort keys: entry L = 4*row + k is (addr column k, value column k) of that row; the last
// n_pub entries are replaced by the public memory (add_pub_memory_in_public_input_section, air.rs:488-506)
__global__ void __launch_bounds__(256) cairo_aux_keys(CairoAux P, unsigned long long* __restrict__ keys, unsigned* __restrict__ idx,
                                                      unsigned short* __restrict__ okeys, unsigned long long* __restrict__ key_or) {
    const unsigned long long L = (unsigned long long)blockIdx.x * 256 + threadIdx.x;
    const unsigned long long total = 4 * P.n;
    unsigned long long k = 0;
    if (L < total) {
        const unsigned long long first_pub = total - P.n_pub;
        k = L >= first_pub ? P.pub_addr[L - first_pub] : fe_low64(ld_fe(P.main + (CAIRO_PC - P.col0 + (L & 3)) * P.n + (L >> 2)));
        keys[L] = k;
        idx[L] = (unsigned)L;
    }
    // OR of all keys: its bit length bounds the number of radix passes
    for (int o = 16; o > 0; o >>= 1) k |= __shfl_xor_sync(0xffffffffu, k, o);
    if ((threadIdx.x & 31) == 0 && k) atomicOr(key_or, k);
    if (L < 3 * P.n) okeys[L] = (unsigned short)fe_low64(ld_fe(P.main + (CAIRO_OFF_DST - P.col0 + L % 3) * P.n + L / 3));
}
// numerators / denominators of the memory permutation argument (air.rs:535-563) and the sorted
// address / value columns
__global__ void __launch_bounds__(256) cairo_aux_terms(CairoAux P, const unsigned* __restrict__ sorted_idx, fe* __restrict__ num,
                                                       fe* __restrict__ den) {
    const unsigned long long L = (unsigned long long)blockIdx.x * 256 + threadIdx.x;
    const unsigned long long total = 4 * P.n, first_pub = total - P.n_pub;
    if (L >= total) return;
    const unsigned long long i = L >> 2, k = L & 3;
    const fe a = ld_fe(P.main + (CAIRO_PC - P.col0 + k) * P.n + i), v = ld_fe(P.main + (CAIRO_INST - P.col0 + k) * P.n + i);
    st_fe(num + L, fe_sub_full(P.z, fe_add_full(a, fe_mul_full(P.alpha, v))));
    const unsigned long long s = sorted_idx[L];
    fe as, vs;
    if (s >= first_pub) {
        as = ld_fe(P.pub_addr_fe + (s - first_pub));
        vs = ld_fe(P.pub_val + (s - first_pub));
    } else {
        as = ld_fe(P.main + (CAIRO_PC - P.col0 + (s & 3)) * P.n + (s >> 2));
        vs = ld_fe(P.main + (CAIRO_INST - P.col0 + (s & 3)) * P.n + (s >> 2));
    }
    st_fe(den + L, fe_sub_full(P.z, fe_add_full(as, fe_mul_full(P.alpha, vs))));
    st_fe(P.aux + (3 + k) * P.n + i, as);
    st_fe(P.aux + (7 + k) * P.n + i, vs);
}
// range-check permutation argument (air.rs:564-588, 683-700)
__global__ void __launch_bounds__(256) cairo_aux_rc_terms(CairoAux P, const unsigned short* __restrict__ sorted_off, fe* __restrict__ num,
                                                          fe* __restrict__ den) {
    const unsigned long long R = (unsigned long long)blockIdx.x * 256 + threadIdx.x;
    if (R >= 3 * P.n) return;
    const unsigned long long i = R / 3, k = R % 3;
    st_fe(num + R, fe_sub_full(P.zrc, ld_fe(P.main + (CAIRO_OFF_DST - P.col0 + k) * P.n + i)));
    fe so = fe_zero();
    so.l[0] = sorted_off[R];
    so = fe_to_mont(so);
    st_fe(den + R, fe_sub_full(P.zrc, so));
    st_fe(P.aux + k * P.n + i, so);
}
// perm[i] = numprefix[i] * densuffix[i+1] / dentotal, de-interleaved into `width` columns starting at aux column `col0`
__global__ void __launch_bounds__(256) cairo_aux_finish(fe* __restrict__ aux, unsigned long long n, unsigned col0, unsigned width,
                                                        const fe* __restrict__ numpre, const fe* __restrict__ densuf, fe inv_total) {
    const unsigned long long L = (unsigned long long)blockIdx.x * 256 + threadIdx.x;
    const unsigned long long total = (unsigned long long)width * n;
    if (L >= total) return;
    fe r = fe_mul_full(ld_fe(numpre + L), inv_total);
    if (L + 1 < total) r = fe_mul_full(r, ld_fe(densuf + L + 1));
    st_fe(aux + (col0 + L % width) * n + L / width, r);
}

// ------------------------------------------------------------------------------------------------
// constraint evaluation over the LDE coset
constexpr int CAIRO_MAX_BOUNDARY = 8;
constexpr int CAIRO_MAX_TRANSITION = 50;
constexpr int CAIRO_EVAL_THREADS = 128;

struct CairoEval {
    // The kernel works on a block of `rows` consecutive LDE rows starting at global row `row0` (the whole
    // coset on one GPU; this rank's row block in a sharded proof).  The "next" frame row i + blowup of the
    // last `blowup` rows lives in the halo (on one GPU: the first rows of the same table, stride m).
    const fe* main;              // column-major LDE of the main trace [main_cols][stride]
    const fe* aux;               // column-major LDE of the auxiliary trace [18][stride]
    const fe* hmain;             // halo: rows row0+rows .. row0+rows+blowup-1 (mod m) of the main columns [main_cols][hstride]
    const fe* haux;
    unsigned long long stride, hstride;
    unsigned long long row0, rows;
    unsigned long long m;        // LDE rows (global)
    unsigned blowup;
    unsigned main_cols;          // 34, or 43 with the range-check builtin
    unsigned has_rc;
    const fe* dom;               // dom[i] = offset * w^i
    const fe* T;                 // T[i] = 1 / (dom[i] - 1)
    fe alpha, z, zrc;            // RAP challenges
    fe g_last;                   // g^(n-1): root of the exemption polynomial
    unsigned nb;
    unsigned bcol[CAIRO_MAX_BOUNDARY];             // column in the combined row (main then aux)
    unsigned long long bshift[CAIRO_MAX_BOUNDARY]; // (blowup * step) mod m
    fe bval[CAIRO_MAX_BOUNDARY];
    // device tables indexed [k * blowup + (i mod blowup)]:
    //   bcoef = g^(-step_k) * (alpha_k * x^n + beta_k)
    //   tcoef = (alpha_k * x^(2n - n (deg_k - 1)) + beta_k) / (x^n - 1)
    const fe* bcoef;
    const fe* tcoef;
    fe two, b15, b16, b32, b48;  // 2, 2^15, 2^16, 2^32, 2^48
    fe* out;                     // [rows]
};

// a load the compiler treats as distinct from every other load of the same address (it is served by L1):
// used to re-read a value instead of keeping it in registers across unrelated work
__device__ __forceinline__ fe ld_fe_again(const fe* p) {
    fe r;
    asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%8];\n\t"
                 "ld.global.v4.u32 {%4, %5, %6, %7}, [%8 + 16];"
                 : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]), "=r"(r.l[6]), "=r"(r.l[7])
                 : "l"(p));
    return r;
}

// exemption flags of CairoAIR::new (air.rs:613-625): constraints that hold everywhere but the last row
__device__ __forceinline__ constexpr bool cairo_exempt(int k) {
    return (k >= 20 && k <= 23) || k == 34 || k == 38 || k == 42 || k == 45;
}

// One thread per LDE row.  Arithmetic is lazy (fe.cuh): products are < 2p, sums/differences carry their
// bound in the comments (in units of p) and stay below 31p, which is what fe_mul accepts against a
// reduced coefficient; the four accumulators are reduced after every addition.  The kernel runs in
// phases that reload what they need (L1/L2 hits) so that few values are live at a time.
#define LM fe_mul
#define LA fe_add_lazy
#define LS1 fe_sub_lazy<1>
#define ACC(dst, a, b) dst = fe_reduce(fe_add_lazy(dst, fe_mul(a, b)))

// PHASE 0: flags, instruction, operand addresses (writes out); 1: register updates and opcodes (adds);
// 2: memory, range check, boundary constraints (adds).  Three launches keep the live set of each small
// (no spills, 4-5 blocks per SM); the extra traffic is two read-modify-writes of `out`.
template <int PHASE>
__global__ void __launch_bounds__(CAIRO_EVAL_THREADS, PHASE == 1 ? 3 : 4) cairo_constraints_kernel(CairoEval P) {
    const unsigned long long il = (unsigned long long)blockIdx.x * CAIRO_EVAL_THREADS + threadIdx.x;   // row inside the block
    if (il >= P.rows) return;
    const unsigned long long i = P.row0 + il;                      // global LDE row
    const unsigned long long n2 = il + P.blowup;                   // Frame::read_from_trace, offsets [0, 1]
    const bool in_block = n2 < P.rows;
    const unsigned r = (unsigned)(i & (P.blowup - 1));
    const fe* tc = P.tcoef + r;
    const unsigned bs = P.blowup;
    const fe* mrow = P.main + il;
    const fe* mnxt = in_block ? P.main + n2 : P.hmain + (n2 - P.rows);
    const fe* arow = P.aux + il;
    const fe* anx = in_block ? P.aux + n2 : P.haux + (n2 - P.rows);
    const unsigned long long m = P.m, cs = P.stride, ns = in_block ? P.stride : P.hstride;
#define CUR(j) ld_fe(mrow + (unsigned long long)(j) * cs)
#define RCUR(j) ld_fe_again(mrow + (unsigned long long)(j) * cs)
#define NXT(j) ld_fe(mnxt + (unsigned long long)(j) * ns)
#define ACUR(j) ld_fe(arow + (unsigned long long)(j) * cs)
#define ANXT(j) ld_fe(anx + (unsigned long long)(j) * ns)
#define COEF(k) ldg_fe(tc + (unsigned)(k) * bs)
    const fe one = fe_one();
    fe acc = fe_zero(), acc_ex = fe_zero();          // plain and exempted (times x - g^(n-1)) constraints
    fe sel = fe_zero(), sel_ex = fe_zero();          // constraints 16..30: times the selector (enforce_selector, air.rs:986-991)

    // ---- phase A: flags (air.rs:869-898), instruction decomposition, operand addresses (air.rs:900-927)
    if constexpr (PHASE == 0) {
        fe f0 = fe_zero();
#pragma unroll 1
        for (int j = 14; j >= 0; --j) {
            const fe f = CUR(j);
            const fe c = LM(f, LS1(f, one));                                   // f (f - 1)             < 2
            ACC(acc, COEF(j), c);
            f0 = fe_reduce(LA(f, LA(f0, f0)));
        }
        ACC(acc, COEF(15), CUR(15));
        const fe off_dst = CUR(27), off_op0 = CUR(28), off_op1 = CUR(29);
        {
            fe s = LA(off_dst, LM(P.b16, off_op0));                              // < 3
            s = LA(s, LM(P.b32, off_op1));                                       // < 5
            s = LA(s, LM(P.b48, f0));                                            // < 7
            ACC(sel, COEF(16), LS1(s, CUR(23)));                                 // INST                  < 8
        }
        const fe ap = CUR(17), fp = CUR(18);
        const fe fpap = LS1(fp, ap);                                             // < 2
        {   // f*fp + (1-f)*ap = ap + f*(fp - ap)
            fe s = LA(LA(ap, LM(CUR(0), fpap)), off_dst);                        // < 4
            ACC(sel, COEF(17), LS1(LS1(s, P.b15), CUR(20)));                     // DST_ADDR              < 6
            s = LA(LA(ap, LM(CUR(1), fpap)), off_op0);
            ACC(sel, COEF(18), LS1(LS1(s, P.b15), CUR(21)));                     // OP0_ADDR              < 6
        }
        {
            const fe f_val = CUR(2), f_fp = CUR(3), f_ap = CUR(4);
            fe s = LA(LA(LM(f_val, CUR(19)), LM(f_ap, ap)), LM(f_fp, fp));       // < 6
            const fe rest = LS1(LS1(LS1(one, f_val), f_ap), f_fp);               // < 4
            s = LA(s, LM(rest, CUR(25)));                                        // < 8
            s = LA(s, off_op1);                                                  // < 9
            ACC(sel, COEF(19), LS1(LS1(s, P.b15), CUR(22)));                     // OP1_ADDR              < 11
        }
        ACC(acc, sel, CUR(33));
    }
    // ---- phase B: register updates (air.rs:929-964) and opcodes (air.rs:966-984).  Values are re-read where
    // they are used (RCUR: a load the compiler may not merge with an earlier one) instead of being kept live.
    if constexpr (PHASE == 1) {
        const fe f_jnz = CUR(9), f_call = CUR(12);
        {
            fe s = LA(LA(RCUR(17), LM(CUR(10), RCUR(16))), CUR(11));             // < 4
            s = LA(s, LA(f_call, f_call));                                       // < 6
            ACC(sel_ex, COEF(20), LS1(s, NXT(17)));                              // NEXT_AP               < 7
        }
        {
            const fe f_ret = CUR(13);
            fe q = LA(LM(f_ret, RCUR(24)), LM(f_call, LA(RCUR(17), P.two)));     // < 4
            q = LA(q, LM(LS1(LS1(one, f_ret), f_call), RCUR(18)));               // < 6
            ACC(sel_ex, COEF(21), LS1(q, NXT(18)));                              // NEXT_FP               < 7
        }
        {
            const fe pc = RCUR(19);
            const fe pc_plus = LA(pc, LA(RCUR(2), one));                         // pc + instruction size  < 3
            const fe pc_next = NXT(19);
            ACC(sel_ex, COEF(22), LM(LS1(RCUR(31), f_jnz), fe_sub_lazy<3>(pc_next, pc_plus)));        // NEXT_PC_1  (2)(4)
            fe lhs = LA(LM(RCUR(30), fe_sub_lazy<2>(pc_next, LA(pc, RCUR(26)))), LM(LS1(one, f_jnz), pc_next));   // < 4
            const fe f_abs = CUR(7), f_rel = CUR(8);
            const fe reg = LS1(LS1(LS1(one, f_abs), f_rel), f_jnz);              // < 4
            const fe res = RCUR(16);
            fe rhs = LA(LM(reg, pc_plus), LM(f_abs, res));                       // (4)(3) ok, < 4
            rhs = LA(rhs, LM(f_rel, LA(pc, res)));                               // < 6
            ACC(sel_ex, COEF(23), fe_sub_lazy<6>(lhs, rhs));                     // NEXT_PC_2             < 10
        }
        {
            const fe t0 = RCUR(30), res = RCUR(16);
            ACC(sel, COEF(24), LS1(LM(f_jnz, RCUR(24)), t0));                    // T0                    < 3
            ACC(sel, COEF(25), LS1(LM(t0, res), RCUR(31)));                      // T1                    < 3
        }
        {
            const fe op0 = RCUR(25), op1 = RCUR(26), mul = CUR(32);
            ACC(sel, COEF(26), fe_sub_lazy<2>(mul, LM(op0, op1)));               // MUL_1                 < 3
            const fe f_add = CUR(5), f_mul = CUR(6);
            fe u = LA(LM(f_add, LA(op0, op1)), LM(f_mul, mul));                  // < 4
            u = LA(u, LM(LS1(LS1(LS1(one, f_add), f_mul), f_jnz), op1));         // < 6
            ACC(sel, COEF(27), fe_sub_lazy<2>(u, LM(LS1(one, f_jnz), RCUR(16)))); // MUL_2                 < 8
        }
        {
            const fe dst = RCUR(24);
            ACC(sel, COEF(28), LM(f_call, LS1(dst, RCUR(18))));                  // CALL_1
            ACC(sel, COEF(30), LM(CUR(14), LS1(dst, RCUR(16))));                 // ASSERT_EQ
            const fe pc_plus = LA(RCUR(19), LA(RCUR(2), one));
            ACC(sel, COEF(29), LM(f_call, fe_sub_lazy<3>(RCUR(25), pc_plus)));   // CALL_2  (1)(4)
        }
        const fe selector = CUR(33);
        ACC(acc, sel, selector);
        ACC(acc_ex, sel_ex, selector);
    }
    // ---- phase C: memory -- increasing addresses, single-valued, permutation argument (air.rs:993-1096)
    if constexpr (PHASE == 2) {
        fe a0 = ACUR(3), v0 = ACUR(7), p0 = ACUR(11);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const fe a1 = k < 3 ? ACUR(3 + k + 1) : ANXT(3), v1 = k < 3 ? ACUR(7 + k + 1) : ANXT(7), p1 = k < 3 ? ACUR(11 + k + 1) : ANXT(11);
            const fe step = LS1(LS1(a1, a0), one);                               // < 3
            const fe c_inc = LM(LS1(a0, a1), step), c_val = LM(LS1(v0, v1), step);
            // original (address, value) of slot k+1; slot 4 is slot 0 of the next row
            const fe ao = k == 0 ? CUR(20) : k == 1 ? CUR(21) : k == 2 ? CUR(22) : NXT(19);
            const fe vo = k == 0 ? CUR(24) : k == 1 ? CUR(25) : k == 2 ? CUR(26) : NXT(23);
            const fe l = LM(fe_sub_lazy<3>(P.z, LA(a1, LM(P.alpha, v1))), p1);    // (4)(1)
            const fe rr = LM(fe_sub_lazy<3>(P.z, LA(ao, LM(P.alpha, vo))), p0);
            const fe c_perm = fe_sub_lazy<2>(l, rr);                             // < 4
            if (k < 3) {
                ACC(acc, COEF(31 + k), c_inc);
                ACC(acc, COEF(35 + k), c_val);
                ACC(acc, COEF(39 + k), c_perm);
            } else {
                ACC(acc_ex, COEF(31 + k), c_inc);
                ACC(acc_ex, COEF(35 + k), c_val);
                ACC(acc_ex, COEF(39 + k), c_perm);
            }
            a0 = a1; v0 = v1; p0 = p1;
        }
    }
    // ---- phase D: range check -- increasing offsets and permutation argument (air.rs:1098-1139)
    if constexpr (PHASE == 2) {
        fe a0 = ACUR(0), p0 = ACUR(15);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const fe a1 = k < 2 ? ACUR(k + 1) : ANXT(0), p1 = k < 2 ? ACUR(15 + k + 1) : ANXT(15);
            const fe c_inc = LM(LS1(a0, a1), LS1(LS1(a1, a0), one));
            const fe oo = k == 0 ? CUR(28) : k == 1 ? CUR(29) : NXT(27);
            const fe c_perm = fe_sub_lazy<2>(LM(LS1(P.zrc, a1), p1), LM(LS1(P.zrc, oo), p0));        // < 4
            if (k < 2) ACC(acc, COEF(43 + k), c_inc); else ACC(acc_ex, COEF(43 + k), c_inc);
            ACC(acc, COEF(46 + k), c_perm);
            a0 = a1; p0 = p1;
        }
    }
    if (PHASE == 2 && P.has_rc) {   // range_check_builtin (air.rs:1141-1160)
        fe s = fe_zero();
#pragma unroll 1
        for (int k = 7; k >= 0; --k) s = fe_reduce(LA(CUR(34 + k), LM(P.b16, s)));
        ACC(acc, COEF(49), LS1(s, CUR(42)));
    }
    if constexpr (PHASE != 0) {
        ACC(acc, acc_ex, LS1(ld_fe(P.dom + i), P.g_last));
        acc = fe_reduce(fe_add_lazy(acc, ld_fe(P.out + il)));
    }
    // ---- boundary constraints (evaluator.rs:58-122): 1/(x - g^s) = g^(-s) * T[i - blowup*s]
    for (unsigned k = 0; PHASE == 2 && k < P.nb; ++k) {
        const unsigned j = P.bcol[k];
        const fe v = j < P.main_cols ? CUR(j) : ACUR(j - P.main_cols);
        const fe zi = ld_fe(P.T + ((i + m - P.bshift[k]) & (m - 1)));
        ACC(acc, LM(zi, ldg_fe(P.bcoef + k * bs + r)), LS1(v, P.bval[k]));        // (2)(2)
    }
    st_fe(P.out + il, acc);
#undef CUR
#undef RCUR
#undef NXT
#undef ACUR
#undef ANXT
#undef COEF
}
#undef LM
#undef LA
#undef LS1
#undef ACC

// H(x) coefficients -> even / odd parts (Polynomial::even_odd_decomposition, prover.rs:252): out[j][k] = h[2k + j];
// *overflow is set when a coefficient of degree >= 2*half is non-zero.
__global__ void __launch_bounds__(256) split_even_odd(const fe* __restrict__ h, unsigned long long m, unsigned long long half,
                                                      fe* __restrict__ out, unsigned* __restrict__ overflow) {
    const unsigned long long i = (unsigned long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= m) return;
    const fe v = ld_fe(h + i);
    if (i < 2 * half) st_fe(out + (i & 1) * half + (i >> 1), v);
    else if (!fe_is_zero(v)) atomicOr(overflow, 1u);
}

}  // namespace s252
