// cairo.cuh -- device kernels of the Cairo stages that sit between the LDE + commitment calls of
// `prove::<CairoAIR>`:
//
//   build_auxiliary_trace          src/cairo/air.rs:660-729    sort by address, permutation columns
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
// auxiliary trace
constexpr unsigned CAIRO_PC = 19, CAIRO_INST = 23, CAIRO_OFF_DST = 27, CAIRO_AUX_COLS = 18;

struct CairoAux {
    const fe* main;                   // column-major main trace [cols][n] (internal format)
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
                                                      unsigned short* __restrict__ okeys) {
    const unsigned long long L = (unsigned long long)blockIdx.x * 256 + threadIdx.x;
    const unsigned long long total = 4 * P.n;
    if (L < total) {
        const unsigned long long first_pub = total - P.n_pub;
        keys[L] = L >= first_pub ? P.pub_addr[L - first_pub] : fe_low64(ld_fe(P.main + (CAIRO_PC + (L & 3)) * P.n + (L >> 2)));
        idx[L] = (unsigned)L;
    }
    if (L < 3 * P.n) okeys[L] = (unsigned short)fe_low64(ld_fe(P.main + (CAIRO_OFF_DST + L % 3) * P.n + L / 3));
}
// numerators / denominators of the memory permutation argument (air.rs:535-563) and the sorted
// address / value columns
__global__ void __launch_bounds__(256) cairo_aux_terms(CairoAux P, const unsigned* __restrict__ sorted_idx, fe* __restrict__ num,
                                                       fe* __restrict__ den) {
    const unsigned long long L = (unsigned long long)blockIdx.x * 256 + threadIdx.x;
    const unsigned long long total = 4 * P.n, first_pub = total - P.n_pub;
    if (L >= total) return;
    const unsigned long long i = L >> 2, k = L & 3;
    const fe a = ld_fe(P.main + (CAIRO_PC + k) * P.n + i), v = ld_fe(P.main + (CAIRO_INST + k) * P.n + i);
    st_fe(num + L, fe_sub_full(P.z, fe_add_full(a, fe_mul_full(P.alpha, v))));
    const unsigned long long s = sorted_idx[L];
    fe as, vs;
    if (s >= first_pub) {
        as = ld_fe(P.pub_addr_fe + (s - first_pub));
        vs = ld_fe(P.pub_val + (s - first_pub));
    } else {
        as = ld_fe(P.main + (CAIRO_PC + (s & 3)) * P.n + (s >> 2));
        vs = ld_fe(P.main + (CAIRO_INST + (s & 3)) * P.n + (s >> 2));
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
    st_fe(num + R, fe_sub_full(P.zrc, ld_fe(P.main + (CAIRO_OFF_DST + k) * P.n + i)));
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
    const fe* main;              // column-major LDE of the main trace [main_cols][m]
    const fe* aux;               // column-major LDE of the auxiliary trace [18][m]
    unsigned long long m;        // LDE rows
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
    fe* out;                     // [m]
};

#define CMUL fe_mul_full
#define CADD fe_add_full
#define CSUB fe_sub_full

// exemption flags of CairoAIR::new (air.rs:613-625): constraints that hold everywhere but the last row
__device__ __forceinline__ bool cairo_exempt(int k) {
    return (k >= 20 && k <= 23) || k == 34 || k == 38 || k == 42 || k == 45;
}

__global__ void __launch_bounds__(CAIRO_EVAL_THREADS) cairo_constraints_kernel(CairoEval P) {
    const unsigned long long i = (unsigned long long)blockIdx.x * CAIRO_EVAL_THREADS + threadIdx.x;
    if (i >= P.m) return;
    const unsigned long long i2 = (i + P.blowup) & (P.m - 1);     // Frame::read_from_trace, offsets [0, 1]
    const unsigned r = (unsigned)(i & (P.blowup - 1));
    const fe* tc = P.tcoef + r;
    const unsigned bs = P.blowup;
    // value of column j of the combined row (main columns, then auxiliary columns)
    auto cur = [&](unsigned j) { return ld_fe(P.main + (unsigned long long)j * P.m + i); };
    auto nxt = [&](unsigned j) { return ld_fe(P.main + (unsigned long long)j * P.m + i2); };
    auto acur = [&](unsigned j) { return ld_fe(P.aux + (unsigned long long)j * P.m + i); };
    auto anxt = [&](unsigned j) { return ld_fe(P.aux + (unsigned long long)j * P.m + i2); };
    auto coef = [&](int k) { return ldg_fe(tc + (unsigned)k * bs); };
    const fe one = fe_one();
    fe acc = fe_zero(), acc_ex = fe_zero();          // plain and exempted (times x - g^(n-1)) constraints
    auto put = [&](int k, const fe& c) {
        const fe t = CMUL(coef(k), c);
        if (cairo_exempt(k)) acc_ex = CADD(acc_ex, t); else acc = CADD(acc, t);
    };

    // ---- flags: bit constraints and f0~ (air.rs:869-898)
    fe f0 = fe_zero();
#pragma unroll 1
    for (int j = 14; j >= 0; --j) {
        const fe f = cur(j);
        put(j, CMUL(f, CSUB(f, one)));
        f0 = CADD(f, CADD(f0, f0));
    }
    put(15, cur(15));

    // ---- constraints 16..30 are multiplied by the selector (enforce_selector, air.rs:986-991)
    fe sel = fe_zero(), sel_ex = fe_zero();
    auto puts = [&](int k, const fe& c) {
        const fe t = CMUL(coef(k), c);
        if (cairo_exempt(k)) sel_ex = CADD(sel_ex, t); else sel = CADD(sel, t);
    };
    const fe ap = cur(17), fp = cur(18), pc = cur(19);
    const fe dst = cur(24), op0 = cur(25), op1 = cur(26), res = cur(16);
    const fe off_dst = cur(27), off_op0 = cur(28), off_op1 = cur(29);
    const fe t0 = cur(30), t1 = cur(31), mul = cur(32);
    fe b15 = fe_zero(), b16 = fe_zero(), b32 = fe_zero(), b48 = fe_zero(), two = fe_zero();
    b15.l[0] = 1u << 15; b16.l[0] = 1u << 16; b32.l[1] = 1u; b48.l[1] = 1u << 16; two.l[0] = 2;
    b15 = fe_to_mont(b15); b16 = fe_to_mont(b16); b32 = fe_to_mont(b32); b48 = fe_to_mont(b48); two = fe_to_mont(two);
    {   // INST
        fe s = CADD(off_dst, CMUL(b16, off_op0));
        s = CADD(s, CMUL(b32, off_op1));
        s = CADD(s, CMUL(b48, f0));
        puts(16, CSUB(s, cur(23)));
    }
    {   // DST_ADDR, OP0_ADDR, OP1_ADDR (air.rs:900-927):  f*fp + (1-f)*ap = ap + f*(fp - ap)
        const fe fpap = CSUB(fp, ap);
        puts(17, CSUB(CADD(CADD(ap, CMUL(cur(0), fpap)), CSUB(off_dst, b15)), cur(20)));
        puts(18, CSUB(CADD(CADD(ap, CMUL(cur(1), fpap)), CSUB(off_op0, b15)), cur(21)));
        const fe f_val = cur(2), f_fp = cur(3), f_ap = cur(4);
        fe s = CADD(CADD(CMUL(f_val, pc), CMUL(f_ap, ap)), CMUL(f_fp, fp));
        s = CADD(s, CMUL(CSUB(CSUB(CSUB(one, f_val), f_ap), f_fp), op0));
        s = CADD(s, CSUB(off_op1, b15));
        puts(19, CSUB(s, cur(22)));
    }
    const fe f_jnz = cur(9), f_call = cur(12), f_ret = cur(13);
    const fe inst_size = CADD(cur(2), one);
    const fe pc_next = nxt(19);
    {   // NEXT_AP, NEXT_FP, NEXT_PC_1, NEXT_PC_2, T0, T1 (air.rs:929-964)
        fe s = CADD(ap, CMUL(cur(10), res));
        s = CADD(s, cur(11));
        s = CADD(s, CADD(f_call, f_call));
        puts(20, CSUB(s, nxt(17)));
        fe q = CADD(CMUL(f_ret, dst), CMUL(f_call, CADD(ap, two)));
        q = CADD(q, CMUL(CSUB(CSUB(one, f_ret), f_call), fp));
        puts(21, CSUB(q, nxt(18)));
        const fe pc_plus = CADD(pc, inst_size);
        puts(22, CMUL(CSUB(t1, f_jnz), CSUB(pc_next, pc_plus)));
        const fe f_abs = cur(7), f_rel = cur(8);
        fe lhs = CADD(CMUL(t0, CSUB(pc_next, CADD(pc, op1))), CMUL(CSUB(one, f_jnz), pc_next));
        fe rhs = CMUL(CSUB(CSUB(CSUB(one, f_abs), f_rel), f_jnz), pc_plus);
        rhs = CADD(rhs, CMUL(f_abs, res));
        rhs = CADD(rhs, CMUL(f_rel, CADD(pc, res)));
        puts(23, CSUB(lhs, rhs));
        puts(24, CSUB(CMUL(f_jnz, dst), t0));
        puts(25, CSUB(CMUL(t0, res), t1));
        // opcode constraints (air.rs:966-984)
        puts(26, CSUB(mul, CMUL(op0, op1)));
        const fe f_add = cur(5), f_mul = cur(6);
        fe u = CADD(CMUL(f_add, CADD(op0, op1)), CMUL(f_mul, mul));
        u = CADD(u, CMUL(CSUB(CSUB(CSUB(one, f_add), f_mul), f_jnz), op1));
        puts(27, CSUB(u, CMUL(CSUB(one, f_jnz), res)));
        puts(28, CMUL(f_call, CSUB(dst, fp)));
        puts(29, CMUL(f_call, CSUB(op0, pc_plus)));
        puts(30, CMUL(cur(14), CSUB(dst, res)));
    }
    {
        const fe selector = cur(33);
        acc = CADD(acc, CMUL(sel, selector));
        acc_ex = CADD(acc_ex, CMUL(sel_ex, selector));
    }

    // ---- memory: increasing addresses, single-valued, permutation argument (air.rs:993-1096)
    {
        const fe a_orig[4] = {nxt(19), cur(20), cur(21), cur(22)};     // a0_next, a1, a2, a3
        const fe v_orig[4] = {nxt(23), dst, op0, op1};
        fe a0 = acur(3), v0 = acur(7), p0 = acur(11);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const fe a1 = k < 3 ? acur(3 + k + 1) : anxt(3), v1 = k < 3 ? acur(7 + k + 1) : anxt(7), p1 = k < 3 ? acur(11 + k + 1) : anxt(11);
            const fe step = CSUB(CSUB(a1, a0), one);
            put(31 + k, CMUL(CSUB(a0, a1), step));
            put(35 + k, CMUL(CSUB(v0, v1), step));
            const int o = (k + 1) & 3;
            const fe l = CMUL(CSUB(P.z, CADD(a1, CMUL(P.alpha, v1))), p1);
            const fe rr = CMUL(CSUB(P.z, CADD(a_orig[o], CMUL(P.alpha, v_orig[o]))), p0);
            put(39 + k, CSUB(l, rr));
            a0 = a1; v0 = v1; p0 = p1;
        }
    }
    // ---- range check: increasing offsets and permutation argument (air.rs:1098-1139)
    {
        const fe o_orig[3] = {nxt(27), off_op0, off_op1};
        fe a0 = acur(0), p0 = acur(15);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const fe a1 = k < 2 ? acur(k + 1) : anxt(0), p1 = k < 2 ? acur(15 + k + 1) : anxt(15);
            put(43 + k, CMUL(CSUB(a0, a1), CSUB(CSUB(a1, a0), one)));
            const int o = (k + 1) % 3;
            put(46 + k, CSUB(CMUL(CSUB(P.zrc, a1), p1), CMUL(CSUB(P.zrc, o_orig[o]), p0)));
            a0 = a1; p0 = p1;
        }
    }
    if (P.has_rc) {   // range_check_builtin (air.rs:1141-1160)
        fe s = fe_zero();
#pragma unroll 1
        for (int k = 7; k >= 0; --k) s = CADD(cur(34 + k), CMUL(b16, s));
        put(49, CSUB(s, cur(42)));
    }
    const fe x = ld_fe(P.dom + i);
    acc = CADD(acc, CMUL(acc_ex, CSUB(x, P.g_last)));

    // ---- boundary constraints (evaluator.rs:58-122): 1/(x - g^s) = g^(-s) * T[i - blowup*s]
    for (unsigned k = 0; k < P.nb; ++k) {
        const unsigned j = P.bcol[k];
        const fe v = j < P.main_cols ? cur(j) : acur(j - P.main_cols);
        const fe zi = ld_fe(P.T + ((i + P.m - P.bshift[k]) & (P.m - 1)));
        acc = CADD(acc, CMUL(CMUL(zi, ldg_fe(P.bcoef + k * bs + r)), CSUB(v, P.bval[k])));
    }
    st_fe(P.out + i, acc);
}
#undef CMUL
#undef CADD
#undef CSUB

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
