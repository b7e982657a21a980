// ntt.cuh -- shared-memory-staged NTT passes over Stark252 (natural order in, natural order out).
//
// A size-N transform (N = 2^n) is run as one, two or three passes ("four-step" decomposition):
//
//   strided pass  (kind A)  view the column as [outer][L][inner]; for every (o, i) run a size-L DIT
//                           NTT along the middle axis, multiply output k by the inter-pass twiddle
//                           ptw[k*inner + i] = s^i * w_(L*inner)^(k*i) [* 1/N] and store in place
//                           of the input position.  A block owns a tile of T consecutive i, so
//                           every global access is a run of T*32 bytes.
//   final pass    (kind B)  rows of L contiguous elements; a block owns T rows whose outputs are
//                           adjacent (consecutive k1), runs a size-L NTT on each and scatters
//                           output k of row (k1,k2) to  (k1 + N1*k2 + N1*N2*k)*ostride + coset.
//
// Coset evaluation (evaluate_offset_fft: out[i] = p(h*w_M^i), M = b*N) is b independent size-N
// coset transforms (coset r has shift s_r = h*w_M^r and lands on outputs b*k + r).  A coset
// transform costs nothing extra: the shift only changes the level twiddles of the FIRST pass
// (level of size n uses s^(L/n... ) * w_n^j) and its inter-pass table.
//
// Inside a block the tile lives in shared memory as [pos][T]; every level is one radix-2 DIT
// butterfly per work item (a, b) -> (a + w*b, a - w*b + 2p) with lazy reduction: values grow by 2p
// per level (< 24p after 11 levels, 32p is the limit), and are brought back to [0,p) once, at the
// store.  The butterfly is ~175 integer instructions, so the 8 shared-memory accesses per level
// are noise; HBM traffic per pass is one read and one write of the data.
#pragma once
#include "fe.cuh"

namespace s252 {

#ifndef S252_NTT_MIN_BLOCKS
#define S252_NTT_MIN_BLOCKS 3          /* resident 64 KB blocks per SM the register budget is sized for */
#endif
#ifndef S252_NTT_NO_F2
#define S252_NTT_NO_F2 0               /* 1: drop the second swizzle term of the tile (fewer address instructions, more bank conflicts) */
#endif
constexpr int NTT_THREADS = 256;
constexpr int NTT_TILE_LOG = 11;                       // 2048 elements = 64 KB of shared memory
constexpr int NTT_TILE = 1 << NTT_TILE_LOG;
constexpr int NTT_MAX_LOGL = NTT_TILE_LOG;

struct NttPass {
    const fe* in;
    fe* out;
    const fe* lvl;        // level twiddles: [ncosets][L]; level of half-size h uses lvl[h + j], j < h
    const fe* ptw;        // kind A: [ncosets][L*inner] inter-pass twiddles
    const fe* oscale;     // kind B, optional: [N1*N2*L] multiplier per natural output index
    unsigned long long in_col_stride, out_col_stride;   // elements between columns
    unsigned long long in_coset_stride, out_coset_stride;   // elements between cosets (0 = shared)
    unsigned logL, logT;
    unsigned logInner;    // kind A
    unsigned logOuter;    // kind A
    unsigned logN1, logN2;   // kind B: digits that precede this one (rows = N1*N2 per column)
    unsigned ncosets;     // b (1 for plain transforms)
    unsigned ostride;     // kind B: output index multiplier (b for coset evaluation, else 1)
    unsigned ncols;
    unsigned lvl_per_coset;   // 1 if the level twiddles differ per coset
    unsigned ptw_per_coset;   // kind A: 1 if the inter-pass twiddles differ per coset
    unsigned rows_are_cols;   // kind B single-pass: the T rows of a tile are T different columns
    unsigned in_lw;           // kind B single-pass: input is in the reference's LW element format
    unsigned out_lw;          // kind B: store outputs in LW format
    unsigned first_unit;      // the level-1 twiddle of this pass is 1 (no coset shift): skip that multiply
    // One transform shared by several GPUs (four-step): a GPU runs only its range of tiles.
    unsigned block0;          // added to blockIdx.x (strided pass: a contiguous tile range = a range of inner positions or of outer rows)
    unsigned part_g0, part_gn;   // kind B: this GPU's range of k1 groups [part_g0, part_g0 + part_gn); part_gn = 0: all of them
};

__device__ __forceinline__ unsigned bitrev32(unsigned x, unsigned bits) { return bits ? __brev(x) >> (32 - bits) : 0u; }

// Shared-memory tile: 2^logE elements (E = L*T), element e = pos*T + t, stored as two planes of
// 16-byte halves (lo limbs 0-3, hi limbs 4-7) so that a quarter-warp touching 8 consecutive elements
// covers all 32 banks, and XOR-swizzled with the top three index bits so that the bit-reversed
// scatter of the load phase (which walks the TOP bits) is conflict-free as well.
struct Tile {
    uint4* lo;
    uint4* hi;
    unsigned swz_shift, swz_mask;      // top three index bits -> bank bits (bit-reversed load scatter)
    unsigned f2_shift, f2_mask, f2_lsh;   // item index bits of the first radix-4 step -> free bank bits
    __device__ __forceinline__ unsigned phys(unsigned e) const {
#if S252_NTT_NO_F2
        return e ^ ((e >> swz_shift) & swz_mask);
#else
        return e ^ ((e >> swz_shift) & swz_mask) ^ (((e >> f2_shift) & f2_mask) << f2_lsh);
#endif
    }
    __device__ __forceinline__ fe ld(unsigned e) const {
        const unsigned q = phys(e);
        const uint4 a = lo[q], b = hi[q];
        return fe{{a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w}};
    }
    __device__ __forceinline__ void st(unsigned e, const fe& v) const {
        const unsigned q = phys(e);
        lo[q] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
        hi[q] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
    }
};
__device__ __forceinline__ Tile make_tile(unsigned char* smem, unsigned logE, unsigned logT) {
    Tile t;
    t.lo = reinterpret_cast<uint4*>(smem);
    t.hi = t.lo + (1u << logE);
    t.swz_shift = logE >= 9 ? logE - 3 : 0;
    t.swz_mask = logE >= 9 ? 7u : 0u;
    // In the first radix-4 step element e = (4j + r)*T + t: the 8 threads of a shared-memory phase
    // differ in t (low logT bits) and in j (bits 2+logT and up), which would all land on the same
    // banks; fold the j bits that sit above bit 2 into the bank bits t leaves free.
    t.f2_shift = 0; t.f2_mask = 0; t.f2_lsh = 0;
    if (logE >= 9 && !S252_NTT_NO_F2) {
        if (logT == 0) { t.f2_shift = 3; t.f2_mask = 3; t.f2_lsh = 0; }
        else if (logT == 1) { t.f2_shift = 3; t.f2_mask = 3; t.f2_lsh = 1; }
        else if (logT == 2) { t.f2_shift = 4; t.f2_mask = 1; t.f2_lsh = 2; }
    }
    return t;
}

// (a, b) -> (a + w*b, a - w*b + 2p); UNIT (compile time) skips the multiplication when w is known to be 1
template <bool UNIT>
__device__ __forceinline__ void bfly(fe& a, fe& b, const fe* __restrict__ tw) {
    fe v;
    if constexpr (UNIT) v = b; else v = fe_mul(b, ldg_fe(tw));
    const fe s = fe_add_lazy(a, v);
    b = fe_sub_lazy<2>(a, v);
    a = s;
}

// One radix-4 step (DIT levels lev and lev+1) over the whole tile: four elements, four butterflies, three
// twiddle loads per work item.  FIRST_UNIT: lev == 0 of a transform without coset shift, where the level-1
// twiddle and the first level-2 twiddle are 1 -- three of the four multiplications disappear at compile time.
// warp_sync: the NEXT step is another radix-4 step whose four inputs per item were all written by items of the same aligned run of
// (4 << lev) << logT <= 32 work items, i.e. by lanes of the same warp (items w = tid + k*256 keep their lane in every step): a warp
// barrier orders those shared-memory accesses and the block barrier is skipped.
#ifndef S252_NTT_WARP_SYNC
#define S252_NTT_WARP_SYNC 1
#endif
template <bool FIRST_UNIT>
__device__ __forceinline__ void radix4_step(const Tile& sm, const fe* __restrict__ lvl, unsigned logL, unsigned logT, unsigned lev,
                                            bool warp_sync = false) {
    const unsigned T = 1u << logT;
    const unsigned half = 1u << lev;
    const unsigned items = (1u << (logL - 2)) << logT;
    for (unsigned w = threadIdx.x; w < items; w += NTT_THREADS) {
        const unsigned t = w & (T - 1);
        const unsigned j = w >> logT;
        const unsigned jj = j & (half - 1);
        const unsigned i0 = ((j >> lev) << (lev + 2)) + jj;
        const unsigned e0 = (i0 << logT) + t, es = half << logT;
        fe x0 = sm.ld(e0), x1 = sm.ld(e0 + es), x2 = sm.ld(e0 + 2 * es), x3 = sm.ld(e0 + 3 * es);
        bfly<FIRST_UNIT>(x0, x1, lvl + half + jj);
        bfly<FIRST_UNIT>(x2, x3, lvl + half + jj);
        bfly<FIRST_UNIT>(x0, x2, lvl + 2 * half + jj);      // level-2 twiddles are {1, w_4}: the first is 1 too
        bfly<false>(x1, x3, lvl + 3 * half + jj);
        sm.st(e0, x0); sm.st(e0 + es, x1); sm.st(e0 + 2 * es, x2); sm.st(e0 + 3 * es, x3);
    }
    if (warp_sync) __syncwarp(); else __syncthreads();
}

// All levels of a size-L DIT transform on every one of the T sequences held in the tile (input
// already in bit-reversed position order).  Levels are taken two at a time in registers, which halves
// the shared-memory traffic and the number of barriers of a level-by-level radix-2 loop.  first_unit: the
// level-1 twiddle is 1 (every transform except the first pass of a coset evaluation).
__device__ __forceinline__ void block_dit(const Tile& sm, const fe* __restrict__ lvl, unsigned logL, unsigned logT,
                                          bool first_unit) {
    const unsigned T = 1u << logT;
    unsigned lev = 0;
    // a step may end in a warp barrier when another radix-4 step follows (lev + 3 < logL) and its producers sit in the same warp
    auto weak = [&](unsigned l) { return S252_NTT_WARP_SYNC && l + 3 < logL && (((4u << l) << logT) <= 32u); };
    if (logL >= 2) {
        if (first_unit) radix4_step<true>(sm, lvl, logL, logT, 0, weak(0));
        else radix4_step<false>(sm, lvl, logL, logT, 0, weak(0));
        lev = 2;
    }
#pragma unroll 1
    for (; lev + 1 < logL; lev += 2) radix4_step<false>(sm, lvl, logL, logT, lev, weak(lev));
    if (lev < logL) {   // odd number of levels: one radix-2 level remains
        const unsigned half = 1u << lev;
        const unsigned items = (1u << (logL - 1)) << logT;
        for (unsigned w = threadIdx.x; w < items; w += NTT_THREADS) {
            const unsigned t = w & (T - 1);
            const unsigned j = w >> logT;
            const unsigned jj = j & (half - 1);
            const unsigned i0 = ((j >> lev) << (lev + 1)) + jj;
            const unsigned e0 = (i0 << logT) + t, es = half << logT;
            fe a = sm.ld(e0), b = sm.ld(e0 + es);
            if (first_unit && lev == 0) bfly<true>(a, b, lvl + half + jj);
            else bfly<false>(a, b, lvl + half + jj);
            sm.st(e0, a); sm.st(e0 + es, b);
        }
        __syncthreads();
    }
}

// ---- kind A: strided pass -------------------------------------------------------------------
__global__ void __launch_bounds__(NTT_THREADS, S252_NTT_MIN_BLOCKS) ntt_pass_strided(NttPass P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Tile sm = make_tile(smem_raw, P.logL + P.logT, P.logT);
    const unsigned L = 1u << P.logL, T = 1u << P.logT;
    // block order: column fastest, then coset, then tile -- the blocks that share an inter-pass
    // twiddle slice (same tile and coset, different columns) run together, so the slice is read from
    // HBM once and served from L2 to the other columns
    const unsigned bidx = blockIdx.x + P.block0;
    const unsigned col = bidx % P.ncols;
    const unsigned rest = bidx / P.ncols;
    const unsigned coset = rest % P.ncosets;
    const unsigned tile = rest / P.ncosets;
    const unsigned tiles_per_outer = 1u << (P.logInner - P.logT);
    const unsigned long long o = tile / tiles_per_outer;
    const unsigned long long i0 = (unsigned long long)(tile % tiles_per_outer) << P.logT;
    const fe* in = P.in + col * P.in_col_stride + coset * P.in_coset_stride + ((o << P.logL) << P.logInner) + i0;
    fe* out = P.out + col * P.out_col_stride + coset * P.out_coset_stride + ((o << P.logL) << P.logInner) + i0;
    const unsigned n = L << P.logT;
#pragma unroll 4
    for (unsigned e = threadIdx.x; e < n; e += NTT_THREADS) {
        const unsigned t = e & (T - 1), pos = e >> P.logT;
        const fe v = ld_fe(in + ((unsigned long long)pos << P.logInner) + t);
        sm.st((bitrev32(pos, P.logL) << P.logT) + t, v);
    }
    __syncthreads();
    block_dit(sm, P.lvl + (P.lvl_per_coset ? (size_t)coset << P.logL : 0), P.logL, P.logT, P.first_unit != 0);
    const fe* ptw = P.ptw + (P.ptw_per_coset ? (size_t)coset << (P.logL + P.logInner) : 0) + i0;
#pragma unroll 2
    for (unsigned e = threadIdx.x; e < n; e += NTT_THREADS) {
        const unsigned t = e & (T - 1), k = e >> P.logT;
        fe v = sm.ld(e);
        const fe w = ldg_fe(ptw + ((unsigned long long)k << P.logInner) + t);
        v = fe_reduce(fe_mul(v, w));
        st_fe(out + ((unsigned long long)k << P.logInner) + t, v);
    }
}

// ---- kind B: final pass ----------------------------------------------------------------------
__global__ void __launch_bounds__(NTT_THREADS, S252_NTT_MIN_BLOCKS) ntt_pass_final(NttPass P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Tile sm = make_tile(smem_raw, P.logL + P.logT, P.logT);
    const unsigned L = 1u << P.logL, T = 1u << P.logT;
    unsigned bid = blockIdx.x, col0 = 0;
    if (!P.rows_are_cols) { col0 = bid % P.ncols; bid /= P.ncols; }
    const unsigned coset = bid % P.ncosets;
    const unsigned tile = bid / P.ncosets;
    const unsigned n = L << P.logT;
    unsigned long long out_base;     // natural output index of (row t = 0, k = 0)
    const fe* in;
    fe* out;
    unsigned live = T;     // live: rows of the tile that exist
    if (P.rows_are_cols) {
        col0 = tile << P.logT;
        live = min(T, P.ncols - col0);
        in = P.in + col0 * P.in_col_stride + coset * P.in_coset_stride;
        out = P.out + col0 * P.out_col_stride;
        out_base = 0;
    } else {
        // tile -> (k2, group of T consecutive k1); a GPU sharing the transform owns a range of the groups
        const unsigned groups = P.part_gn ? P.part_gn : 1u << (P.logN1 - P.logT);
        const unsigned k2 = tile / groups;
        const unsigned k1 = (P.part_g0 + tile % groups) << P.logT;
        in = P.in + col0 * P.in_col_stride + coset * P.in_coset_stride +
             ((((unsigned long long)k1 << P.logN2) + k2) << P.logL);
        out = P.out + col0 * P.out_col_stride;
        out_base = k1 + ((unsigned long long)k2 << P.logN1);
    }
    // row t of the tile starts at in + t*row_stride
    const unsigned long long row_stride = P.rows_are_cols ? P.in_col_stride : (1ull << (P.logN2 + P.logL));
#pragma unroll 4
    for (unsigned e = threadIdx.x; e < n; e += NTT_THREADS) {
        const unsigned pos = e & (L - 1), t = e >> P.logL;
        fe v = fe_zero();
        if (t < live) {
            const fe* src = in + t * row_stride + pos;
            v = P.in_lw ? ld_lw(src) : ld_fe(src);
        }
        sm.st((bitrev32(pos, P.logL) << P.logT) + t, v);
    }
    __syncthreads();
    block_dit(sm, P.lvl + (P.lvl_per_coset ? (size_t)coset << P.logL : 0), P.logL, P.logT, P.first_unit != 0);
    const unsigned logRows = P.logN1 + P.logN2;
    for (unsigned e = threadIdx.x; e < n; e += NTT_THREADS) {
        const unsigned t = e & (T - 1), k = e >> P.logT;
        if (t >= live) continue;
        fe v = sm.ld(e);
        unsigned long long oi;     // natural output index within the column
        fe* dst;
        if (P.rows_are_cols) {
            oi = k;
            dst = out + t * P.out_col_stride + (unsigned long long)k * P.ostride + coset;
        } else {
            oi = out_base + t + ((unsigned long long)k << logRows);
            dst = out + oi * P.ostride + coset;
        }
        if (P.oscale) v = fe_mul(v, ldg_fe(P.oscale + oi));
        v = fe_reduce(v);
        if (P.out_lw) st_lw(dst, v); else st_fe(dst, v);
    }
}

// ---- table generation ---------------------------------------------------------------------
// lvl[c][h + j] = shift_c^(L/(2h)) * w_(2h)^j  for h = 1, 2, .., L/2, j < h   (lvl[c][0] unused)
// where shift_c = base_shift * coset_step^c  and  w_(2h) = wL^(L/(2h)).
__global__ void gen_level_twiddles(fe* lvl, unsigned logL, unsigned ncosets, fe wL, fe base_shift, fe coset_step) {
    const unsigned L = 1u << logL;
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= L * ncosets) return;
    const unsigned c = idx >> logL, x = idx & (L - 1);
    if (x == 0) { st_fe(lvl + idx, fe_one()); return; }
    const unsigned lev = 31 - __clz(x);            // h = 2^lev
    const unsigned j = x - (1u << lev);
    const unsigned e = L >> (lev + 1);             // L/(2h)
    const fe shift = fe_mul_full(base_shift, fe_pow(coset_step, c));
    const fe a = fe_pow(shift, e);
    const fe b = fe_pow(wL, (uint64_t)e * j);
    st_fe(lvl + idx, fe_mul_full(a, b));
}
// ptw[c][k*inner + i] = scale * (shift_c * wS^k)^i,  S = L*inner, k < L, i < inner
__global__ void gen_pass_twiddles(fe* ptw, unsigned logL, unsigned logInner, unsigned ncosets, fe wS, fe base_shift,
                                  fe coset_step, fe scale) {
    const unsigned long long S = 1ull << (logL + logInner);
    const unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= S * ncosets) return;
    const unsigned c = (unsigned)(idx >> (logL + logInner));
    const unsigned long long x = idx & (S - 1);
    const unsigned long long k = x >> logInner, i = x & ((1ull << logInner) - 1);
    const fe shift = fe_mul_full(base_shift, fe_pow(coset_step, c));
    const fe g = fe_mul_full(shift, fe_pow(wS, k));
    st_fe(ptw + idx, fe_mul_full(scale, fe_pow(g, i)));
}
// out[k] = scale * base^k
__global__ void gen_powers(fe* out, unsigned long long n, fe base, fe scale) {
    const unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    st_fe(out + idx, fe_mul_full(scale, fe_pow(base, idx)));
}

}  // namespace s252
