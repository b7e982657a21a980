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

// the same with L1-bypassing loads: for children written by OTHER blocks of the running kernel
__device__ __forceinline__ void hash_pair_cg(const uint64_t* children, uint64_t out[4]) {
    uint64_t st[25];
#pragma unroll
    for (int k = 0; k < 8; ++k) st[k] = __ldcg(reinterpret_cast<const unsigned long long*>(children) + k);
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

// ---- device-side Fiat-Shamir chain of the FRI commit phase ---------------------------------------------
// fri_commit_phase (fri/mod.rs:33-69) alternates  transcript.append(root_k)  and  zeta_(k+1) = transcript_to_field():
// every fold needs the root of the layer before it.  Keeping the DefaultTranscript sponge on the device lets the
// whole phase be queued without a host round trip per layer; the host replays the same appends on its own
// transcript afterwards from the roots it reads back once.
constexpr int FRI_MAX_LAYERS = 64;
struct FriChain {
    uint64_t sponge[25];          // DefaultTranscript's Keccak state (bytes already absorbed are XORed in)
    unsigned fill;                // bytes absorbed into the current block
    unsigned layer;               // index of the next root to be appended
    fe c;                         // zeta_k / (2 h_(k-1)) for the next fold (Montgomery form)
    fe zeta;                      // the last sampled zeta (Montgomery form)
    fe last_value;                // fri_last_value (Montgomery form), written by the tail
    fe inv2h[FRI_MAX_LAYERS];     // 1 / (2 h^(2^k))
    uint64_t roots[FRI_MAX_LAYERS][4];
};
__device__ inline void chain_absorb32(FriChain* ch, const uint64_t w[4]) {      // 32 bytes, memory order = lane byte order
    unsigned fill = ch->fill;
    for (int i = 0; i < 32; ++i) {
        const uint64_t byte = (w[i >> 3] >> (8 * (i & 7))) & 0xff;
        ch->sponge[fill >> 3] ^= byte << (8 * (fill & 7));
        if (++fill == 136) {
            uint64_t st[25];
#pragma unroll
            for (int k = 0; k < 25; ++k) st[k] = ch->sponge[k];
            keccak_f1600(st);
#pragma unroll
            for (int k = 0; k < 25; ++k) ch->sponge[k] = st[k];
            fill = 0;
        }
    }
    ch->fill = fill;
}
__device__ __forceinline__ uint64_t bswap64(uint64_t x) {
    return ((uint64_t)bswap32((uint32_t)x) << 32) | bswap32((uint32_t)(x >> 32));
}
// DefaultTranscript::challenge + transcript_to_field (transcript.rs:23-43): finalize, reverse the digest, reseed the
// sponge with it; the field element is the reversed digest read big-endian with the top 5 bits cleared, i.e. the
// digest itself read as a little-endian 256-bit integer below 2^251.  Returns it in Montgomery form.
__device__ inline fe chain_to_field(FriChain* ch) {
    uint64_t st[25];
#pragma unroll
    for (int k = 0; k < 25; ++k) st[k] = ch->sponge[k];
    const unsigned fill = ch->fill;
    // padding byte at a run-time position: select the lane without dynamic register indexing
#pragma unroll
    for (int k = 0; k < 17; ++k) if ((fill >> 3) == (unsigned)k) st[k] ^= (uint64_t)0x01 << (8 * (fill & 7));
    st[16] ^= 0x8000000000000000ULL;
    keccak_f1600(st);
    fe v;
#pragma unroll
    for (int k = 0; k < 4; ++k) { v.l[2 * k] = (uint32_t)st[k]; v.l[2 * k + 1] = (uint32_t)(st[k] >> 32); }
    v.l[7] &= 0x07ffffffu;
#pragma unroll
    for (int k = 0; k < 25; ++k) ch->sponge[k] = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) ch->sponge[k] = bswap64(st[3 - k]);
    ch->fill = 32;
    return fe_to_mont(v);
}
// transcript.append(root_k); zeta_(k+1) = transcript_to_field()   (fri/mod.rs:37,41,54): one thread.
__device__ inline void chain_advance(FriChain* ch, const uint64_t root[4]) {
    const unsigned k = ch->layer;
    chain_absorb32(ch, root);
#pragma unroll
    for (int j = 0; j < 4; ++j) ch->roots[k][j] = root[j];
    const fe z = chain_to_field(ch);
    ch->zeta = z;
    ch->c = fe_reduce(fe_mul(z, ch->inv2h[k]));
    ch->layer = k + 1;
}

// The top of a tree in ONE launch: block b hashes children [512 b, 512 b + 512) of `child_level` up to their
// common ancestor (nine levels through shared memory, as many as the level has); the block that finishes last
// (ticket) then hashes the remaining levels up to the root, and -- for a FRI layer -- advances the transcript
// chain.  child_level <= 18.  nodes: heap array of 4 x u64 digests; *ticket must be 0 and is left 0.
constexpr int MERKLE_BLOCK = 256;
constexpr int MERKLE_FUSED_LEVELS = 9;
constexpr unsigned MERKLE_FINISH_MAX_LEVEL = 2 * MERKLE_FUSED_LEVELS;
template <bool CG>
__device__ __forceinline__ void block_reduce_levels(uint64_t* __restrict__ nodes, uint64_t* sm, unsigned level,
                                                    unsigned long long first, unsigned levels) {
    unsigned active = 1u << (levels - 1);
    for (unsigned d = 0; d < levels; ++d) {
        const unsigned long long pfirst = first >> 1;
        uint64_t out[4];
        if (threadIdx.x < active) {
            const uint64_t* ch = nodes + 4 * (((1ull << level) - 1) + first + 2 * threadIdx.x);
            if (d != 0) hash_pair(sm + 8 * threadIdx.x, out);
            else if (CG) hash_pair_cg(ch, out);
            else hash_pair(ch, out);
        }
        __syncthreads();
        if (threadIdx.x < active) {
            store_digest(nodes + 4 * (((1ull << (level - 1)) - 1) + pfirst + threadIdx.x), out);
            store_digest(sm + 4 * threadIdx.x, out);
        }
        __syncthreads();
        first = pfirst;
        level -= 1;
        active >>= 1;
    }
}
__global__ void __launch_bounds__(MERKLE_BLOCK) merkle_finish(uint64_t* __restrict__ nodes, unsigned child_level,
                                                             unsigned* __restrict__ ticket, FriChain* __restrict__ chain) {
    __shared__ __align__(16) uint64_t sm[MERKLE_BLOCK * 4];
    __shared__ unsigned last;
    const unsigned lv0 = min(child_level, (unsigned)MERKLE_FUSED_LEVELS);
    if (lv0) block_reduce_levels<false>(nodes, sm, child_level, (unsigned long long)blockIdx.x << lv0, lv0);
    unsigned level = child_level - lv0;
    if (gridDim.x > 1) {
        if (threadIdx.x == 0) {
            __threadfence();
            last = atomicAdd(ticket, 1u) == gridDim.x - 1;
        }
        __syncthreads();
        if (!last) return;
        __threadfence();
        if (threadIdx.x == 0) *ticket = 0;
        if (level) block_reduce_levels<true>(nodes, sm, level, 0, level);
    }
    if (chain && threadIdx.x == 0) {
        uint64_t root[4];
        if (child_level) { root[0] = sm[0]; root[1] = sm[1]; root[2] = sm[2]; root[3] = sm[3]; }
        else { root[0] = nodes[0]; root[1] = nodes[1]; root[2] = nodes[2]; root[3] = nodes[3]; }
        chain_advance(chain, root);
    }
}

// ---- FRI: evaluation-domain fold fused with the next layer's leaf hashing ------------------
// The reference folds coefficients and re-evaluates every layer (fri/mod.rs:43-54); the values
// are those of  out[i] = (v+s)/2 + zeta*(v-s)/(2*x_i),  v = layer[i], s = layer[i+size/2],
// x_i = h_k*w^i  -- the verifier's own formula (verifier.rs:511-512).
// inv_tw[j*tw_stride] = w_size^(-j);  c = zeta/(2*h_k);  inv2 = 1/2 (all Montgomery form).
__device__ __forceinline__ fe fri_fold_value(const fe& v, const fe& s, const fe& tw, const fe& c, const fe& inv2) {
    const fe w = fe_mul(tw, c);                                        // < 2p
    const fe sum = fe_add_lazy(v, s);                                  // < 2p
    const fe dif = fe_sub_lazy<1>(v, s);                               // < 2p
    return fe_reduce(fe_add_lazy(fe_mul(sum, inv2), fe_mul(dif, w)));
}
__device__ __forceinline__ void fri_leaf_digest(const fe& r, uint64_t d[4]) {       // Keccak256(r.to_bytes_be())
    uint64_t st[25];
#pragma unroll
    for (int k = 0; k < 25; ++k) st[k] = 0;
    uint64_t wds[4];
    fe_be_lanes(fe_from_mont(r), wds);
    st[0] = wds[0]; st[1] = wds[1]; st[2] = wds[2]; st[3] = wds[3];
    st[4] = 0x01ULL;
    st[16] = 0x8000000000000000ULL;
    keccak_f1600(st);
    d[0] = st[0]; d[1] = st[1]; d[2] = st[2]; d[3] = st[3];
}
__global__ void __launch_bounds__(128) fri_fold_commit(const fe* __restrict__ layer, unsigned long long half,
                                                       const fe* __restrict__ inv_tw, unsigned long long tw_stride,
                                                       fe c_imm, const FriChain* __restrict__ chain, fe inv2,
                                                       fe* __restrict__ out, uint64_t* __restrict__ leaves) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= half) return;
    const fe c = chain ? ld_fe(&chain->c) : c_imm;                     // zeta / (2 h_k): sampled on the device or by the host
    const fe r = fri_fold_value(ld_fe(layer + i), ld_fe(layer + i + half), ldg_fe(inv_tw + i * tw_stride), c, inv2);
    st_fe(out + i, r);
    if (leaves) {
        uint64_t d[4];
        fri_leaf_digest(r, d);
        store_digest(leaves + 4 * i, d);
    }
}

// The fold on a block of rows of the next layer, operands given separately (a rank of a sharded FRI: v and s arrive
// from the two ranks that hold layer[i] and layer[i + size/2]): out[j] = fold(v[j], s[j]) for row i0 + j.
__global__ void __launch_bounds__(128) fri_fold_rows_kernel(const fe* __restrict__ v, const fe* __restrict__ s, unsigned long long count,
                                                            const fe* __restrict__ inv_tw, unsigned long long tw_stride,
                                                            unsigned long long i0, fe c, fe inv2, fe* __restrict__ out) {
    const unsigned long long j = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    st_fe(out + j, fri_fold_value(ld_fe(v + j), ld_fe(s + j), ldg_fe(inv_tw + (i0 + j) * tw_stride), c, inv2));
}

// The tail of the commit phase in ONE block: once a layer has at most FRI_TAIL_MAX evaluations, every remaining
// fold, leaf hash, tree, transcript step and the final  last_value = mean of the last evaluations  (fri/mod.rs:43-69)
// run here back to back -- a launch + root read-back per layer costs more than these layers' arithmetic.
constexpr unsigned FRI_TAIL_LOG_MAX = 14;
constexpr int FRI_TAIL_THREADS = 512;
struct FriTail {
    FriChain* chain;
    const fe* in;                 // evaluations of the last layer committed before the tail (size = 2^log_size)
    unsigned log_size;
    unsigned n_commit;            // folds whose result is a committed layer; one more, uncommitted, fold follows
    fe* evals[FRI_TAIL_LOG_MAX + 1];        // outputs of fold f (f = n_commit: the uncommitted remainder)
    uint64_t* nodes[FRI_TAIL_LOG_MAX + 1];
    const fe* inv_tw;
    unsigned long long tw_stride;           // of the first tail fold; doubles with every fold
    fe inv2, inv_last;                      // 1/2 and 1/(number of evaluations left after the last fold), Montgomery form
};
__global__ void __launch_bounds__(FRI_TAIL_THREADS) fri_tail_kernel(FriTail P) {
    __shared__ __align__(16) fe red[FRI_TAIL_THREADS];
    const fe* in = P.in;
    unsigned logs = P.log_size;
    unsigned long long stride = P.tw_stride;
    for (unsigned f = 0; f <= P.n_commit; ++f) {
        const unsigned half = 1u << (logs - 1);
        const bool commit = f < P.n_commit;
        const fe c = ld_fe(&P.chain->c);
        fe* out = P.evals[f];
        uint64_t* nodes = P.nodes[f];
        for (unsigned i = threadIdx.x; i < half; i += FRI_TAIL_THREADS) {
            const fe r = fri_fold_value(ld_fe(in + i), ld_fe(in + i + half), ldg_fe(P.inv_tw + i * stride), c, P.inv2);
            st_fe(out + i, r);
            if (commit) {
                uint64_t d[4];
                fri_leaf_digest(r, d);
                store_digest(nodes + 4 * ((unsigned long long)half - 1 + i), d);
            }
        }
        __syncthreads();
        if (commit) {
            for (unsigned lev = logs - 1; lev > 0; --lev) {
                const unsigned parents = 1u << (lev - 1);
                for (unsigned q = threadIdx.x; q < parents; q += FRI_TAIL_THREADS) {
                    uint64_t d[4];
                    hash_pair(nodes + 4 * (((1ull << lev) - 1) + 2 * q), d);
                    store_digest(nodes + 4 * (((1ull << (lev - 1)) - 1) + q), d);
                }
                __syncthreads();
            }
            if (threadIdx.x == 0) {
                const uint64_t root[4] = {nodes[0], nodes[1], nodes[2], nodes[3]};
                chain_advance(P.chain, root);
            }
            __syncthreads();
        }
        in = out;
        logs -= 1;
        stride <<= 1;
    }
    // last_value = coefficient 0 of the fully folded polynomial = mean of its evaluations on the remaining coset
    const unsigned left = 1u << logs;
    fe acc = fe_zero();
    for (unsigned i = threadIdx.x; i < left; i += FRI_TAIL_THREADS) acc = fe_reduce(fe_add_lazy(acc, ld_fe(in + i)));
    red[threadIdx.x] = acc;
    __syncthreads();
    for (unsigned s = FRI_TAIL_THREADS / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] = fe_reduce(fe_add_lazy(red[threadIdx.x], red[threadIdx.x + s]));
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const fe lv = fe_reduce(fe_mul(red[0], P.inv_last));
        st_fe(&P.chain->last_value, lv);
        uint64_t w[4];
        fe_be_lanes(fe_from_mont(lv), w);
        chain_absorb32(P.chain, w);                                   // transcript.append(last_value), fri/mod.rs:69
    }
}

// ---- grinding (src/starks/grinding.rs:17-48) ---------------------------------------------
// Keccak256(challenge || nonce_le), head = first 8 digest bytes big-endian, accept when
// trailing_zeros(head) >= factor.  The SMALLEST accepted nonce wins (the reference searches 0, 1, 2, ..).
__device__ __forceinline__ bool grind_accepts(uint64_t c0, uint64_t c1, uint64_t c2, uint64_t c3, uint64_t nonce, unsigned factor) {
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
    return tz >= factor;
}
// One launch for the whole search: the grid walks the nonces [base, limit) in batches of gridDim*blockDim and a
// thread leaves only when its batch starts above the best nonce found so far, so every nonce below the reported
// one has been tested by the time the grid drains: the result is the minimum, as in the sequential search.
// `batches` bounds the launch (the host relaunches from where it stopped if nothing was found).
// Several GPUs share one search by taking the batches round-robin: GPU `part` of `parts` tests batches
// part, part + parts, ..; the minimum over the GPUs' results is the global minimum (a GPU only skips batches
// that lie above a hit of its own).
__global__ void __launch_bounds__(256) grind_kernel(uint64_t c0, uint64_t c1, uint64_t c2, uint64_t c3, uint64_t base,
                                                   uint64_t limit, unsigned batches, unsigned part, unsigned parts,
                                                   unsigned factor, unsigned long long* __restrict__ best) {
    const unsigned long long span = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned b = 0; b < batches; ++b) {
        const unsigned long long off = ((unsigned long long)b * parts + part) * span;
        const uint64_t start = base + off;
        if (start >= limit || start < base) break;                               // end of range / wrapped around 2^64
        if (*(volatile unsigned long long*)best < start) break;
        const uint64_t nonce = start + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
        if (nonce >= start && nonce < limit && grind_accepts(c0, c1, c2, c3, nonce, factor)) atomicMin(best, (unsigned long long)nonce);
    }
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
// The same for a group of `gcols` columns starting at rows[0] of a row-major table whose rows are `row_pitch`
// elements apart, written for a source in PINNED HOST memory: the threads read the caller's table over PCIe
// themselves (consecutive threads -> consecutive elements of a row, UNROLL independent 32-byte loads in flight per
// thread), so the upload, TraceTable::cols() and the format change are one pass and no staging copy exists.
// PCIe (~55 GB/s) is 20x slower than the scattered 32-byte sector writes, so no shared-memory tile is needed;
// a small grid (it only has to keep ~1 MB of reads in flight) leaves the SMs to the transforms of the
// previous column group running on the compute stream.
__global__ void __launch_bounds__(256) rows_lw_to_cols_stream(const fe* __restrict__ rows, unsigned long long nrows,
                                                              unsigned long long row_pitch, unsigned gcols,
                                                              fe* __restrict__ cols, unsigned long long col_stride) {
    constexpr int UNROLL = 4;
    const unsigned long long total = nrows * gcols;
    const unsigned long long nthreads = (unsigned long long)gridDim.x * blockDim.x;
    unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; idx + (UNROLL - 1) * nthreads < total; idx += UNROLL * nthreads) {
        fe v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const unsigned long long e = idx + u * nthreads;
            v[u] = ld_lw(rows + (e / gcols) * row_pitch + (e % gcols));
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const unsigned long long e = idx + u * nthreads;
            st_fe(cols + (e % gcols) * col_stride + (e / gcols), v[u]);
        }
    }
    for (; idx < total; idx += nthreads)
        st_fe(cols + (idx % gcols) * col_stride + (idx / gcols), ld_lw(rows + (idx / gcols) * row_pitch + (idx % gcols)));
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
