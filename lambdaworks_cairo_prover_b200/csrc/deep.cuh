// deep.cuh -- out-of-domain evaluations and the DEEP composition polynomial on the GPU
// (first "next" row).
//
// Reference (CPU, coefficient form):
//   Frame::get_trace_evaluations             src/starks/frame.rs:67-83      Horner per polynomial and point
//   compute_deep_composition_poly            src/starks/prover.rs:410-482   cols*offsets Ruffini divisions
// Here p0 is produced directly as evaluations on the LDE coset from the resident LDE columns,
//   p0(x) = sum_k [ sum_j g_jk * (t_j(x) - t_j(z g^k)) ] / (x - z g^k)
//         + [ g (H1(x) - H1(z^2)) + g' (H2(x) - H2(z^2)) ] / (x - z^2)
// (the verifier's own formula, src/starks/verifier.rs:526-557): same polynomial, same values, and
// FRI layer 0 no longer needs an NTT.
#pragma once
#include "fe.cuh"

namespace s252 {

constexpr int EVAL_THREADS = 256;

// partial[(j * splits + s) * npoints + p] = sum over i = r (mod 256*splits), r = s*256 + t, of coeffs[j][i] * xs[p]^i
// (block s of column j).  Every thread runs one Horner chain per point in y = x^(256*splits) over its
// residue class -- the chains of the different points share the coefficient loads and give the
// scheduler independent multiplications -- then scales by x^r; the block adds its 256 partial sums.
// The host adds the `splits` partials of a (column, point).
constexpr int EVAL_MAX_POINTS = 4;
template <int NP>
__global__ void __launch_bounds__(EVAL_THREADS) poly_eval_points(const fe* __restrict__ coeffs, unsigned long long col_stride,
                                                                  unsigned long long n, const fe* __restrict__ xs,
                                                                  fe* __restrict__ partial, unsigned splits) {
    __shared__ fe part[EVAL_THREADS];
    const unsigned j = blockIdx.x, s = blockIdx.y, t = threadIdx.x;
    const unsigned long long stride = (unsigned long long)EVAL_THREADS * splits;
    const unsigned long long r = (unsigned long long)s * EVAL_THREADS + t;
    const fe* c = coeffs + (unsigned long long)j * col_stride;
    fe x[NP], y[NP], acc[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        x[p] = ld_fe(xs + p);
        acc[p] = fe_zero();
        // y = x^stride (stride is a power of two times 256)
        y[p] = x[p];
        for (unsigned long long e = stride; e > 1; e >>= 1) y[p] = fe_mul_full(y[p], y[p]);
    }
    if (r < n) {
        unsigned long long i = r + ((n - 1 - r) / stride) * stride;    // largest index = r (mod stride) below n
        for (;;) {
            const fe ci = ld_fe(c + i);
#pragma unroll
            for (int p = 0; p < NP; ++p) acc[p] = fe_add_lazy(fe_mul(acc[p], y[p]), ci);      // < 3p
            if (i < stride) break;
            i -= stride;
        }
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            fe xr = fe_one(), b = x[p];                  // x^r
            for (unsigned long long e = r; e; e >>= 1) {
                if (e & 1) xr = fe_mul_full(xr, b);
                b = fe_mul_full(b, b);
            }
            acc[p] = fe_reduce(fe_mul(acc[p], xr));
        }
    }
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        part[t] = acc[p];
        __syncthreads();
        for (unsigned w = EVAL_THREADS / 2; w > 0; w >>= 1) {
            if (t < w) part[t] = fe_add_full(part[t], part[t + w]);
            __syncthreads();
        }
        if (t == 0) st_fe(partial + ((unsigned long long)j * splits + s) * NP + p, part[0]);
        __syncthreads();
    }
}

constexpr int DEEP_THREADS = 128;
constexpr int DEEP_MAX_TABLES = 4;
constexpr int DEEP_MAX_K = 4;

struct DeepParams {
    const fe* cols[DEEP_MAX_TABLES];             // column-major LDE tables: trace tables first, composition table last
    unsigned long long strides[DEEP_MAX_TABLES];
    unsigned ncols[DEEP_MAX_TABLES];
    unsigned ntables;                            // the last one is the composition table (H1, H2)
    unsigned K;                                  // frame rows (transition offsets)
    const fe* gammas;                            // device: [trace_cols][K] trace-term coefficients, then gamma, gamma'
    // 1/(x_i - z g^k) = g^(-k) * U[i - blowup*k]  with  U[i] = 1/(x_i - z):  x_(i - blowup*k) = x_i / g^k
    const fe* U;                                 // 1/(x_i - z) for i = u_base, u_base + 1, .. (mod m): the whole coset, or a block + its halo
    const fe* V;                                 // 1/(x_i - z^2) for i = v_base, ..
    unsigned long long u_base, v_base;           // global row of U[0] / V[0] (0 when the tables cover the whole coset)
    unsigned long long rot[DEEP_MAX_K];          // (blowup * offset_k) mod m
    fe ginv[DEEP_MAX_K];                         // g^(-offset_k)
    fe ck[DEEP_MAX_K];                           // sum_j gamma_jk * t_j(z g^k)
    fe cz2;                                      // gamma*H1(z^2) + gamma'*H2(z^2)
    unsigned long long m;                        // LDE rows (global)
    unsigned long long row0, rows;               // the block of rows this launch covers (0, m on one GPU); cols/out are block-local
    fe* out;                                     // [rows] evaluations of p0 (internal format)
};

// a^(p-2)
__device__ inline fe fe_inverse(const fe& a) {
    fe r = fe_one(), base = a;
    for (int bit = 0; bit < 252; ++bit) {
        const bool set = bit < 192 || bit == 196 || bit == 251;
        if (set) r = fe_mul_full(r, base);
        base = fe_mul_full(base, base);
    }
    return r;
}

// One thread per LDE row: K running sums over all trace columns (lazy: a product adds < 2p, reduced every
// 8 columns), then the K + 1 divisions as multiplications by the precomputed inverse tables.
template <int K>
__global__ void __launch_bounds__(DEEP_THREADS) deep_composition_kernel(DeepParams P) {
    const unsigned long long row = (unsigned long long)blockIdx.x * DEEP_THREADS + threadIdx.x;    // row inside the block
    if (row >= P.rows) return;
    const unsigned long long gi = P.row0 + row;                                                     // global LDE row
    fe s[K];
#pragma unroll
    for (int k = 0; k < K; ++k) s[k] = fe_zero();
    const fe* g = P.gammas;
    unsigned pending = 0;
    for (unsigned tb = 0; tb + 1 < P.ntables; ++tb) {
        const fe* col = P.cols[tb] + row;
        for (unsigned j = 0; j < P.ncols[tb]; ++j) {
            const fe v = ld_fe(col + (unsigned long long)j * P.strides[tb]);
#pragma unroll
            for (int k = 0; k < K; ++k) s[k] = fe_add_lazy(s[k], fe_mul(v, ldg_fe(g + k)));        // + < 2p
            g += K;
            if (++pending == 8) {
#pragma unroll
                for (int k = 0; k < K; ++k) s[k] = fe_reduce(s[k]);                                  // < 17p before
                pending = 0;
            }
        }
    }
    const unsigned ct = P.ntables - 1;
    const fe h1 = ld_fe(P.cols[ct] + row), h2 = ld_fe(P.cols[ct] + P.strides[ct] + row);
    const fe sz = fe_reduce(fe_add_lazy(fe_mul(h1, ldg_fe(g)), fe_mul(h2, ldg_fe(g + 1))));
    fe acc = fe_mul_full(fe_sub_full(sz, P.cz2), ld_fe(P.V + (gi - P.v_base)));
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const fe inv = fe_mul(P.ginv[k], ld_fe(P.U + ((gi + 2 * P.m - P.rot[k] - P.u_base) & (P.m - 1))));   // < 2p
        acc = fe_reduce(fe_add_lazy(acc, fe_mul(fe_sub_lazy<1>(fe_reduce(s[k]), P.ck[k]), inv)));    // (2)(2)
    }
    st_fe(P.out + row, acc);
}

}  // namespace s252
