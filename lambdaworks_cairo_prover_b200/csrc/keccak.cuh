// keccak.cuh -- Keccak-f[1600] / Keccak-256 (0x01 padding, rate 136) on the integer pipes.
// The reference hashes with the sha3 crate's Keccak256 (Cargo.toml:17): Merkle leaves and nodes
// (src/starks/config.rs:10-20), the Fiat-Shamir transcript and grinding (src/starks/grinding.rs).
// One thread per hash; the 25 lanes stay in registers (50 x u32), rotations are funnel shifts,
// chi is one LOP3 per 32-bit half.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace s252 {

__constant__ uint64_t KECCAK_RC[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
    0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
    0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
    0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
    0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};

__constant__ uint32_t KECCAK_POW2[32] = {
    1u << 0, 1u << 1, 1u << 2, 1u << 3, 1u << 4, 1u << 5, 1u << 6, 1u << 7, 1u << 8, 1u << 9, 1u << 10, 1u << 11,
    1u << 12, 1u << 13, 1u << 14, 1u << 15, 1u << 16, 1u << 17, 1u << 18, 1u << 19, 1u << 20, 1u << 21, 1u << 22,
    1u << 23, 1u << 24, 1u << 25, 1u << 26, 1u << 27, 1u << 28, 1u << 29, 1u << 30, 1u << 31};

// 64-bit rotate by a compile-time amount.  Two flavours share the work between the two integer
// pipes: the funnel-shift form is 2 SHF on the ALU pipe (where all the LOP3s of theta/chi already
// run); the multiply form is 3 IMADs on the FMA pipe, which Keccak otherwise leaves idle:
//   lo*2^n = {lo >> (32-n) : lo << n}   (IMAD.WIDE),   new_hi = hi*2^n + that.hi   (IMAD),
//   new_lo = hi32(hi*2^n) + that.lo     (IMAD.HI)      -- the added bit ranges are disjoint.
#ifndef S252_KECCAK_MUL_ROT
#define S252_KECCAK_MUL_ROT 0          /* how many of the 24 lane rotations per round use the multiply form */
#endif
#ifndef S252_KECCAK_POW2_FROM_CONST
#define S252_KECCAK_POW2_FROM_CONST 0
#endif
__device__ __forceinline__ uint64_t rol64_shf(uint64_t x, int n) {
    uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    if (n == 0) return x;
    if (n == 32) return ((uint64_t)lo << 32) | hi;
    uint32_t rlo, rhi;
    if (n < 32) {
        rlo = __funnelshift_l(hi, lo, n);
        rhi = __funnelshift_l(lo, hi, n);
    } else {
        rlo = __funnelshift_l(lo, hi, n - 32);
        rhi = __funnelshift_l(hi, lo, n - 32);
    }
    return ((uint64_t)rhi << 32) | rlo;
}
__device__ __forceinline__ uint64_t rol64_mul(uint64_t x, int n) {
    uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    if (n == 0) return x;
    if (n == 32) return ((uint64_t)lo << 32) | hi;
    if (n > 32) { const uint32_t t = lo; lo = hi; hi = t; n -= 32; }   // rotate by 32 first (free)
    // the power of two comes from the constant bank so that ptxas keeps real multiplies (with an
    // immediate it strength-reduces the high half back into an ALU-pipe LEA.HI)
    const uint32_t m = S252_KECCAK_POW2_FROM_CONST ? KECCAK_POW2[n] : (1u << n);
    uint32_t plo, phi, rlo, rhi;
    uint64_t pw;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(pw) : "r"(lo), "r"(m));
    plo = (uint32_t)pw;
    phi = (uint32_t)(pw >> 32);
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(rhi) : "r"(hi), "r"(m), "r"(phi));
    asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(rlo) : "r"(hi), "r"(m), "r"(plo));
    return ((uint64_t)rhi << 32) | rlo;
}
// n is a compile-time constant at every call site; `fma` picks the pipe
__device__ __forceinline__ uint64_t rol64(uint64_t x, int n) { return rol64_shf(x, n); }
// lane rotation number `idx` (0..23) of a round
#define S252_ROTL(idx, x, n) (((idx) < S252_KECCAK_MUL_ROT) ? rol64_mul((x), (n)) : rol64_shf((x), (n)))
__device__ __forceinline__ uint64_t xor3(uint64_t a, uint64_t b, uint64_t c) {
    uint32_t lo, hi;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(lo) : "r"((uint32_t)a), "r"((uint32_t)b), "r"((uint32_t)c));
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(hi) : "r"((uint32_t)(a >> 32)), "r"((uint32_t)(b >> 32)), "r"((uint32_t)(c >> 32)));
    return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint64_t xor5(uint64_t a, uint64_t b, uint64_t c, uint64_t d, uint64_t e) {
    return xor3(xor3(a, b, c), d, e);
}
__device__ __forceinline__ uint64_t chi(uint64_t a, uint64_t b, uint64_t c) {
    // a ^ (~b & c): truth table 0xD2 with (a, b, c) as the three LOP3 inputs
    uint32_t lo, hi;
    asm("lop3.b32 %0, %1, %2, %3, 0xD2;" : "=r"(lo) : "r"((uint32_t)a), "r"((uint32_t)b), "r"((uint32_t)c));
    asm("lop3.b32 %0, %1, %2, %3, 0xD2;" : "=r"(hi) : "r"((uint32_t)(a >> 32)), "r"((uint32_t)(b >> 32)), "r"((uint32_t)(c >> 32)));
    return ((uint64_t)hi << 32) | lo;
}

__device__ __forceinline__ void keccak_f1600(uint64_t s[25]) {
#pragma unroll 1
    for (int r = 0; r < 24; ++r) {
        uint64_t C0 = xor5(s[0], s[5], s[10], s[15], s[20]);
        uint64_t C1 = xor5(s[1], s[6], s[11], s[16], s[21]);
        uint64_t C2 = xor5(s[2], s[7], s[12], s[17], s[22]);
        uint64_t C3 = xor5(s[3], s[8], s[13], s[18], s[23]);
        uint64_t C4 = xor5(s[4], s[9], s[14], s[19], s[24]);
        // theta folded into the rho/pi load: s ^ D[x] = s ^ C[x-1] ^ rol(C[x+1], 1) is ONE 3-input LOP3
        // per 32-bit half (the five rotated columns R are shared by the 25 lanes).
        const uint64_t R0 = rol64(C0, 1), R1 = rol64(C1, 1), R2 = rol64(C2, 1), R3 = rol64(C3, 1), R4 = rol64(C4, 1);
#define S252_TH(i, Ca, Rb) xor3(s[i], Ca, Rb)
        // theta + rho + pi: B[y][(2x+3y)%5] = rol(s[x][y] ^ D[x], r[x][y]),  D[x] = C[x-1] ^ R[x+1]
        uint64_t B00 = S252_TH(0, C4, R1);
        uint64_t B10 = S252_ROTL(0, S252_TH(1, C0, R2), 1), B20 = S252_ROTL(1, S252_TH(2, C1, R3), 62), B05 = S252_ROTL(2, S252_TH(3, C2, R4), 28),
                 B15 = S252_ROTL(3, S252_TH(4, C3, R0), 27);
        uint64_t B16 = S252_ROTL(4, S252_TH(5, C4, R1), 36), B01 = S252_ROTL(5, S252_TH(6, C0, R2), 44), B11 = S252_ROTL(6, S252_TH(7, C1, R3), 6),
                 B21 = S252_ROTL(7, S252_TH(8, C2, R4), 55), B06 = S252_ROTL(8, S252_TH(9, C3, R0), 20);
        uint64_t B07 = S252_ROTL(9, S252_TH(10, C4, R1), 3), B17 = S252_ROTL(10, S252_TH(11, C0, R2), 10), B02 = S252_ROTL(11, S252_TH(12, C1, R3), 43),
                 B12 = S252_ROTL(12, S252_TH(13, C2, R4), 25), B22 = S252_ROTL(13, S252_TH(14, C3, R0), 39);
        uint64_t B23 = S252_ROTL(14, S252_TH(15, C4, R1), 41), B08 = S252_ROTL(15, S252_TH(16, C0, R2), 45), B18 = S252_ROTL(16, S252_TH(17, C1, R3), 15),
                 B03 = S252_ROTL(17, S252_TH(18, C2, R4), 21), B13 = S252_ROTL(18, S252_TH(19, C3, R0), 8);
        uint64_t B14 = S252_ROTL(19, S252_TH(20, C4, R1), 18), B24 = S252_ROTL(20, S252_TH(21, C0, R2), 2), B09 = S252_ROTL(21, S252_TH(22, C1, R3), 61),
                 B19 = S252_ROTL(22, S252_TH(23, C2, R4), 56), B04 = S252_ROTL(23, S252_TH(24, C3, R0), 14);
#undef S252_TH
        // names above are B<index> with index = x' + 5*y' already; rows of five:
        // row 0: B00 B01 B02 B03 B04 ; row 1: B05 B06 B07 B08 B09 ; row 2: B10..B14 ; row 3: B15..B19 ; row 4: B20..B24
        s[0] = chi(B00, B01, B02) ^ KECCAK_RC[r];
        s[1] = chi(B01, B02, B03); s[2] = chi(B02, B03, B04); s[3] = chi(B03, B04, B00); s[4] = chi(B04, B00, B01);
        s[5] = chi(B05, B06, B07); s[6] = chi(B06, B07, B08); s[7] = chi(B07, B08, B09); s[8] = chi(B08, B09, B05);
        s[9] = chi(B09, B05, B06);
        s[10] = chi(B10, B11, B12); s[11] = chi(B11, B12, B13); s[12] = chi(B12, B13, B14); s[13] = chi(B13, B14, B10);
        s[14] = chi(B14, B10, B11);
        s[15] = chi(B15, B16, B17); s[16] = chi(B16, B17, B18); s[17] = chi(B17, B18, B19); s[18] = chi(B18, B19, B15);
        s[19] = chi(B19, B15, B16);
        s[20] = chi(B20, B21, B22); s[21] = chi(B21, B22, B23); s[22] = chi(B22, B23, B24); s[23] = chi(B23, B24, B20);
        s[24] = chi(B24, B20, B21);
    }
}

__device__ __forceinline__ uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

}  // namespace s252
