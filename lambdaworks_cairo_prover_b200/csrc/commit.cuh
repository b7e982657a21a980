// commit.cuh -- Merkle commitment, FRI folding, grinding, layout and opening kernels.
//
// Reference semantics (all fixture-verified against the golden proofs):
//   BatchedMerkleTree leaf  = Keccak256(row[0].to_bytes_be() || row[1].to_bytes_be() || ..)
//   FriMerkleTree leaf      = Keccak256(elem.to_bytes_be())
//   node                    = Keccak256(left || right), heap layout, root at 0, leaf i at n-1+i
// to_bytes_be is the 32-byte big-endian CANONICAL value, so every leaf kernel first takes the
// element out of Montgomery form (one sparse reduction, no full multiply).
#pragma once
#include "fe.cuh"
#include "keccak.cuh"

namespace s252 {

// The four Keccak lanes (little-endian u64) that hold the big-endian bytes of canonical c.
__device__ __forceinline__ void fe_be_lanes(const fe& c, uint64_t w[4]) {
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const int m = 3 - s;   // 64-bit chunk m of the value, most significant first
        w[s] = ((uint64_t)bswap32(c.l[2 * m]) << 32) | bswap32(c.l[2 * m + 1]);
    }
}

// Absorb element number jj (0..16) of a 17-element group: lanes (4*jj + s) % 17, permuting when a
// 136-byte block fills.  17 elements are exactly 4 blocks, so positions are compile-time.
template <int JJ>
__device__ __forceinline__ void absorb_elem(uint64_t st[25], const fe& mont) {
    uint64_t w[4];
    fe_be_lanes(fe_from_mont(mont), w);
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        constexpr int base = 4 * JJ;
        const int lane = (base + s) % 17;
        st[lane] ^= w[s];
        if (lane == 16) keccak_f1600(st);
    }
}
template <int JJ>
__device__ __forceinline__ void absorb_tail(uint64_t st[25], const fe* __restrict__ col, unsigned long long col_stride,
                                            unsigned rem) {
    if constexpr (JJ < 17) {
        if (JJ < rem) {
            absorb_elem<JJ>(st, ld_fe(col + (unsigned long long)JJ * col_stride));
            absorb_tail<JJ + 1>(st, col, col_stride, rem);
        } else {
            st[(4 * JJ) % 17] ^= 0x01ULL;   // first padding byte right after the message
        }
    }
}

// One thread per row.  cols: column-major [ncols][col_stride] internal elements; digest of row i
// goes to leaves[i] (32 bytes).
__global__ void __launch_bounds__(128) merkle_leaves(const fe* __restrict__ cols, unsigned long long col_stride,
                                                     unsigned ncols, unsigned long long nrows,
                                                     uint64_t* __restrict__ leaves) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows) return;
    uint64_t st[25];
#pragma unroll
    for (int k = 0; k < 25; ++k) st[k] = 0;
    const fe* col = cols + i;
    unsigned j = 0;
    for (; j + 17 <= ncols; j += 17) {
        const fe* c = col + (unsigned long long)j * col_stride;
        absorb_elem<0>(st, ld_fe(c));
        absorb_elem<1>(st, ld_fe(c + 1 * col_stride));
        absorb_elem<2>(st, ld_fe(c + 2 * col_stride));
        absorb_elem<3>(st, ld_fe(c + 3 * col_stride));
        absorb_elem<4>(st, ld_fe(c + 4 * col_stride));
        absorb_elem<5>(st, ld_fe(c + 5 * col_stride));
        absorb_elem<6>(st, ld_fe(c + 6 * col_stride));
        absorb_elem<7>(st, ld_fe(c + 7 * col_stride));
        absorb_elem<8>(st, ld_fe(c + 8 * col_stride));
        absorb_elem<9>(st, ld_fe(c + 9 * col_stride));
        absorb_elem<10>(st, ld_fe(c + 10 * col_stride));
        absorb_elem<11>(st, ld_fe(c + 11 * col_stride));
        absorb_elem<12>(st, ld_fe(c + 12 * col_stride));
        absorb_elem<13>(st, ld_fe(c + 13 * col_stride));
        absorb_elem<14>(st, ld_fe(c + 14 * col_stride));
        absorb_elem<15>(st, ld_fe(c + 15 * col_stride));
        absorb_elem<16>(st, ld_fe(c + 16 * col_stride));
    }
    absorb_tail<0>(st, col + (unsigned long long)j * col_stride, col_stride, ncols - j);
    st[16] ^= 0x8000000000000000ULL;
    keccak_f1600(st);
    uint64_t* d = leaves + 4 * i;
    d[0] = st[0]; d[1] = st[1]; d[2] = st[2]; d[3] = st[3];
}

__device__ __forceinline__ void hash_pair(const uint64_t* __restrict__ children, uint64_t out[4]) {
    uint64_t st[25];
#pragma unroll
    for (int k = 0; k < 8; ++k) st[k] = children[k];
#pragma unroll
    for (int k = 8; k < 25; ++k) st[k] = 0;
    st[8] ^= 0x01ULL;
    st[16] ^= 0x8000000000000000ULL;
    keccak_f1600(st);
    out[0] = st[0]; out[1] = st[1]; out[2] = st[2]; out[3] = st[3];
}

__device__ __forceinline__ void hash_two(const uint64_t l[4], const uint64_t r[4], uint64_t out[4]) {
    uint64_t st[25];
#pragma unroll
    for (int k = 0; k < 4; ++k) { st[k] = l[k]; st[4 + k] = r[k]; }
#pragma unroll
    for (int k = 8; k < 25; ++k) st[k] = 0;
    st[8] ^= 0x01ULL;
    st[16] ^= 0x8000000000000000ULL;
    keccak_f1600(st);
    out[0] = st[0]; out[1] = st[1]; out[2] = st[2]; out[3] = st[3];
}
__device__ __forceinline__ void store_digest(uint64_t* dst, const uint64_t d[4]) {
    ulonglong2* q = reinterpret_cast<ulonglong2*>(dst);
    q[0] = make_ulonglong2(d[0], d[1]);
    q[1] = make_ulonglong2(d[2], d[3]);
}
// Wide levels: one thread builds a whole LV-level subtree (2^LV children -> 2^LV - 1 ancestors), so
// every thread of the launch stays busy on every level (a per-level halving scheme idles 3/4 of
// its warps).  LV levels per launch; digests of all intermediate levels are written to the heap.
template <int LV>
__global__ void __launch_bounds__(128) merkle_subtrees(uint64_t* __restrict__ nodes, unsigned child_level) {
    const unsigned long long q = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= (1ull << (child_level - LV))) return;
    const uint64_t* ch = nodes + 4 * (((1ull << child_level) - 1) + (q << LV));
    uint64_t* l1 = nodes + 4 * (((1ull << (child_level - 1)) - 1) + (q << (LV - 1)));
    if constexpr (LV == 1) {
        uint64_t p[4];
        hash_pair(ch, p);
        store_digest(l1, p);
    } else if constexpr (LV == 2) {
        uint64_t p0[4], p1[4], g[4];
        hash_pair(ch, p0);
        hash_pair(ch + 8, p1);
        store_digest(l1, p0);
        store_digest(l1 + 4, p1);
        hash_two(p0, p1, g);
        store_digest(nodes + 4 * (((1ull << (child_level - 2)) - 1) + q), g);
    } else {
        uint64_t* l2 = nodes + 4 * (((1ull << (child_level - 2)) - 1) + (q << 1));
        uint64_t g0[4], g1[4], t[4];
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            uint64_t p0[4], p1[4];
            hash_pair(ch + 16 * half, p0);
            hash_pair(ch + 16 * half + 8, p1);
            store_digest(l1 + 8 * half, p0);
            store_digest(l1 + 8 * half + 4, p1);
            if (half == 0) hash_two(p0, p1, g0); else hash_two(p0, p1, g1);
        }
        store_digest(l2, g0);
        store_digest(l2 + 4, g1);
        hash_two(g0, g1, t);
        store_digest(nodes + 4 * (((1ull << (child_level - 3)) - 1) + q), t);
    }
}

// Builds up to MERKLE_FUSED_LEVELS levels above `child_level` (the level whose 2^child_level
// digests already exist).  nodes: heap array of 4 x u64 digests.  Block b owns children
// [b*2*BLOCK, (b+1)*2*BLOCK) of child_level and every ancestor that lies entirely above them.
constexpr int MERKLE_BLOCK = 256;
constexpr int MERKLE_FUSED_LEVELS = 9;
__global__ void __launch_bounds__(MERKLE_BLOCK) merkle_nodes(uint64_t* __restrict__ nodes, unsigned child_level,
                                                            unsigned levels) {
    __shared__ uint64_t sm[MERKLE_BLOCK * 4];
    const unsigned long long nchildren = 1ull << child_level;
    unsigned long long first = (unsigned long long)blockIdx.x * (2 * MERKLE_BLOCK);   // first child index
    unsigned active = (unsigned)min((unsigned long long)MERKLE_BLOCK, nchildren / 2);
    unsigned level = child_level;
    for (unsigned d = 0; d < levels; ++d) {
        // parents at level-1; this block's parents are [first/2, first/2 + active)
        const unsigned long long pfirst = first >> 1;
        uint64_t out[4];
        if (threadIdx.x < active) {
            if (d == 0) {
                const uint64_t* ch = nodes + 4 * (((1ull << level) - 1) + first + 2 * threadIdx.x);
                hash_pair(ch, out);
            } else {
                hash_pair(sm + 8 * threadIdx.x, out);
            }
        }
        __syncthreads();
        if (threadIdx.x < active) {
            uint64_t* dst = nodes + 4 * (((1ull << (level - 1)) - 1) + pfirst + threadIdx.x);
            dst[0] = out[0]; dst[1] = out[1]; dst[2] = out[2]; dst[3] = out[3];
            uint64_t* s = sm + 4 * threadIdx.x;
            s[0] = out[0]; s[1] = out[1]; s[2] = out[2]; s[3] = out[3];
        }
        __syncthreads();
        first = pfirst;
        level -= 1;
        active >>= 1;
        if (active == 0) break;
    }
}

// ---- FRI: evaluation-domain fold fused with the next layer's leaf hashing ------------------
// The reference folds coefficients and re-evaluates every layer (fri/mod.rs:43-54); the values
// are those of  out[i] = (v+s)/2 + zeta*(v-s)/(2*x_i),  v = layer[i], s = layer[i+size/2],
// x_i = h_k*w^i  -- the verifier's own formula (verifier.rs:511-512).
// inv_tw[j*tw_stride] = w_size^(-j);  c = zeta/(2*h_k);  inv2 = 1/2 (all Montgomery form).
__global__ void __launch_bounds__(128) fri_fold_commit(const fe* __restrict__ layer, unsigned long long half,
                                                       const fe* __restrict__ inv_tw, unsigned long long tw_stride,
                                                       fe c, fe inv2, fe* __restrict__ out,
                                                       uint64_t* __restrict__ leaves) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= half) return;
    const fe v = ld_fe(layer + i), s = ld_fe(layer + i + half);
    const fe w = fe_mul(ldg_fe(inv_tw + i * tw_stride), c);            // < 2p
    const fe sum = fe_add_lazy(v, s);                                  // < 2p
    const fe dif = fe_sub_lazy<1>(v, s);                               // < 2p
    const fe r = fe_reduce(fe_add_lazy(fe_mul(sum, inv2), fe_mul(dif, w)));
    st_fe(out + i, r);
    if (leaves) {
        uint64_t st[25];
#pragma unroll
        for (int k = 0; k < 25; ++k) st[k] = 0;
        uint64_t wds[4];
        fe_be_lanes(fe_from_mont(r), wds);
        st[0] = wds[0]; st[1] = wds[1]; st[2] = wds[2]; st[3] = wds[3];
        st[4] = 0x01ULL;
        st[16] = 0x8000000000000000ULL;
        keccak_f1600(st);
        uint64_t* d = leaves + 4 * i;
        d[0] = st[0]; d[1] = st[1]; d[2] = st[2]; d[3] = st[3];
    }
}

// ---- grinding (src/starks/grinding.rs:17-48) ---------------------------------------------
// Thread t tests nonce base + t: Keccak256(challenge || nonce_le), head = first 8 digest bytes
// big-endian, accept when trailing_zeros(head) >= factor.  The SMALLEST accepted nonce wins.
__global__ void __launch_bounds__(256) grind_kernel(uint64_t c0, uint64_t c1, uint64_t c2, uint64_t c3, uint64_t base,
                                                   unsigned long long count, unsigned factor,
                                                   unsigned long long* __restrict__ best) {
    const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const uint64_t nonce = base + t;
    uint64_t st[25];
#pragma unroll
    for (int k = 0; k < 25; ++k) st[k] = 0;
    st[0] = c0; st[1] = c1; st[2] = c2; st[3] = c3;
    st[4] = nonce;                 // to_le_bytes == the lane's own byte order
    st[5] = 0x01ULL;
    st[16] = 0x8000000000000000ULL;
    keccak_f1600(st);
    // head = from_be_bytes(digest[0..8]) = bswap64(lane 0)
    const uint32_t lo = (uint32_t)st[0], hi = (uint32_t)(st[0] >> 32);
    const uint64_t head = ((uint64_t)bswap32(lo) << 32) | bswap32(hi);
    const unsigned tz = head ? (unsigned)(__ffsll((long long)head) - 1) : 64u;
    if (tz >= factor) atomicMin(best, (unsigned long long)nonce);
}

// ---- layout -------------------------------------------------------------------------------
// Row-major table in the reference's LW element format -> column-major internal format.
// 32x32 element tiles through shared memory so both sides move whole 1 KB runs where they can.
__global__ void __launch_bounds__(256) rows_lw_to_cols(const fe* __restrict__ rows, unsigned long long nrows,
                                                       unsigned ncols, fe* __restrict__ cols,
                                                       unsigned long long col_stride) {
    __shared__ uint4 tile[32][33][2];
    const unsigned long long r0 = (unsigned long long)blockIdx.x * 32;
    const unsigned c0 = blockIdx.y * 32;
    for (unsigned e = threadIdx.x; e < 1024; e += 256) {
        const unsigned c = e & 31, r = e >> 5;
        if (r0 + r < nrows && c0 + c < ncols) {
            const fe v = ld_lw(rows + (r0 + r) * ncols + c0 + c);
            tile[r][c][0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
            tile[r][c][1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
        }
    }
    __syncthreads();
    for (unsigned e = threadIdx.x; e < 1024; e += 256) {
        const unsigned r = e & 31, c = e >> 5;
        if (r0 + r < nrows && c0 + c < ncols) {
            uint4* dst = reinterpret_cast<uint4*>(cols + (unsigned long long)(c0 + c) * col_stride + r0 + r);
            dst[0] = tile[r][c][0];
            dst[1] = tile[r][c][1];
        }
    }
}
// element-wise format conversion (same shape)
__global__ void lw_to_internal(const fe* __restrict__ in, fe* __restrict__ out, unsigned long long n) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) st_fe(out + i, ld_lw(in + i));
}
__global__ void internal_to_lw(const fe* __restrict__ in, fe* __restrict__ out, unsigned long long n) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) st_lw(out + i, ld_fe(in + i));
}
// out[i] = in[i*step]   (evaluate_polynomial_on_lde_domain's step rule, prover.rs:118-122)
__global__ void subsample(const fe* __restrict__ in, fe* __restrict__ out, unsigned long long n, unsigned long long step) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) st_fe(out + i, ld_fe(in + i * step));
}

// ---- openings -------------------------------------------------------------------------------
// rows_out[q][j] (LW) = cols[j][idx[q]]
__global__ void gather_rows(const fe* __restrict__ cols, unsigned long long col_stride, unsigned ncols,
                            const unsigned long long* __restrict__ idx, unsigned nq, fe* __restrict__ rows_out) {
    const unsigned e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nq * ncols) return;
    const unsigned q = e / ncols, j = e % ncols;
    st_lw(rows_out + e, ld_fe(cols + (unsigned long long)j * col_stride + idx[q]));
}
// paths_out[q][k] = sibling digest at height k on the way from leaf idx[q] to the root
__global__ void gather_paths(const uint64_t* __restrict__ nodes, unsigned depth,
                             const unsigned long long* __restrict__ idx, unsigned nq, uint64_t* __restrict__ paths_out) {
    const unsigned e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nq * depth) return;
    const unsigned q = e / depth, k = e % depth;
    unsigned long long node = idx[q] + (1ull << depth) - 1;
    for (unsigned s = 0; s < k; ++s) node = (node - 1) >> 1;
    const unsigned long long sib = (node & 1) ? node + 1 : node - 1;
    const uint64_t* src = nodes + 4 * sib;
    uint64_t* dst = paths_out + 4 * (unsigned long long)e;
    dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
}

// fri_query_phase's reads for ALL layers in one launch (fri/mod.rs:74-127).  Layer k has size
// domain >> k; entry e = ((sym * nq + q) * L + k): value of evaluation[(iota_q + sym*size/2) % size] and
// its authentication path, padded to `stride` digests.
struct FriLayerRef { const fe* evals; const uint64_t* nodes; };
__global__ void fri_query_gather(const FriLayerRef* __restrict__ layers, unsigned L, unsigned log_domain,
                                 const unsigned long long* __restrict__ iotas, unsigned nq, unsigned stride,
                                 fe* __restrict__ vals_out, uint64_t* __restrict__ paths_out) {
    const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned per = stride + 1;                       // slot 0: the value, slots 1..stride: path entries
    const unsigned long long total = 2ull * nq * L * per;
    if (t >= total) return;
    const unsigned slot = (unsigned)(t % per);
    const unsigned long long e = t / per;
    const unsigned k = (unsigned)(e % L), q = (unsigned)((e / L) % nq), sym = (unsigned)(e / ((unsigned long long)L * nq));
    const unsigned depth = log_domain - k;
    const unsigned long long size = 1ull << depth;
    const unsigned long long idx = (iotas[q] + (sym ? size / 2 : 0)) & (size - 1);
    const FriLayerRef lay = layers[k];
    if (slot == 0) {
        st_lw(vals_out + e, ld_fe(lay.evals + idx));
        return;
    }
    const unsigned lvl = slot - 1;
    if (lvl >= depth) return;
    unsigned long long node = idx + size - 1;
    for (unsigned s = 0; s < lvl; ++s) node = (node - 1) >> 1;
    const unsigned long long sib = (node & 1) ? node + 1 : node - 1;
    const uint64_t* src = lay.nodes + 4 * sib;
    uint64_t* dst = paths_out + 4 * (e * stride + lvl);
    dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
}

}  // namespace s252
