// fe.cuh -- Stark252 field arithmetic for sm_100a integer pipes.
//
// p = 2^251 + 17*2^192 + 1.  Elements are 8 x u32 little-endian limbs in Montgomery form
// (R = 2^256), the same value the reference's FieldElement<Stark252PrimeField> holds
// (lambdaworks-math montgomery_backed_prime_fields) -- only the limb order
// differs (the reference stores 4 x u64 most-significant first; see ld_lw/st_lw).
//
// Arithmetic is LAZY inside kernels: values live in [0, 32p) (p < 2^252 leaves 4 spare bits),
// fe_mul returns a value < 2p, and only stores that leave a kernel are brought to [0, p).
//
// fe_mul is 80 32x32->64 multiply-adds (64 schoolbook + 16 for the sparse Montgomery reduction) plus carry
// glue.  The reduction uses p = 1 (mod 2^192): with mu = -1 the Montgomery quotient of the low
// 192 bits is just their negation, and  m*p = m + m*(2^59+17)*2^192  needs two small multipliers.  The
// multiplier 2^27 (limb 7 of p) is applied as funnel shifts on the ALU pipe (S252_MONT_SHIFT), which leaves
// 72 IMAD.WIDE.U32 for the 4-cycle multiplier pipe: measured 1.6 % on the NTT passes.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace s252 {

struct fe { uint32_t l[8]; };

#define S252_P0 0x00000001u
#define S252_P6 0x00000011u
#define S252_P7 0x08000000u

__host__ __device__ constexpr fe fe_zero() { return fe{{0, 0, 0, 0, 0, 0, 0, 0}}; }
// R mod p  (Montgomery form of 1)
__host__ __device__ constexpr fe fe_one() {
    return fe{{0xffffffe1u, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xfffffdf0u, 0x07ffffffu}};
}
// R^2 mod p
__host__ __device__ constexpr fe fe_r2() {
    return fe{{0x7e000401u, 0xfffffd73u, 0x330fffffu, 0x00000001u, 0xff6f8000u, 0xffffffffu, 0x5e008810u, 0x07ffd4abu}};
}

// acc[0..7] += {a[0], a[2], a[4], a[6]} * b as four 64-bit products on consecutive limb pairs;
// the carry out is added to acc[8].
__device__ __forceinline__ void mad_row4(uint32_t* acc, const uint32_t* a, uint32_t b) {
    asm("mad.lo.cc.u32  %0, %9,  %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9,  %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32       %8, %8, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
          "+r"(acc[7]), "+r"(acc[8])
        : "r"(a[0]), "r"(a[2]), "r"(a[4]), "r"(a[6]), "r"(b));
}
// same without the carry capture (used where the carry is provably zero)
__device__ __forceinline__ void mad_row4_nc(uint32_t* acc, const uint32_t* a, uint32_t b) {
    asm("mad.lo.cc.u32  %0, %8,  %12, %0;\n\t"
        "madc.hi.cc.u32 %1, %8,  %12, %1;\n\t"
        "madc.lo.cc.u32 %2, %9,  %12, %2;\n\t"
        "madc.hi.cc.u32 %3, %9,  %12, %3;\n\t"
        "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
        "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
        "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
        "madc.hi.cc.u32 %7, %11, %12, %7;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
          "+r"(acc[7])
        : "r"(a[0]), "r"(a[2]), "r"(a[4]), "r"(a[6]), "r"(b));
}

// Montgomery reduction of a 512-bit value T (16 limbs) that ALREADY CONTAINS the constants
//   2^192*(2^59+17+1)  (limbs 6,7: +0x12, +0x08000000)   the "+p" and "+1" of the 192-bit step
//   2^256              (limb 8: +1)                        the "+1" of the 64-bit step
//   2^384*(2^59+17)    (limbs 12,13)                      its "+p"
// Returns (T_without_constants + multiples of p) / 2^256, a value < 2p when T < 31 p^2.
#ifndef S252_MONT_SHIFT
#define S252_MONT_SHIFT 1              /* the 2^27 rows of the reduction as funnel shifts: 72 instead of 80 wide multiplies (0: all 80 as IMAD.WIDE) */
#endif
__device__ __forceinline__ fe mont_reduce(uint32_t T[16]) {
    // Reduction step 1 (192-bit digit).  With n = ~T[0..5]:  T + (n+1)*p  has its low 192 bits
    // equal to 2^192 exactly (carry 1 -> preloaded into limb 6) and gains (n+1)*(2^59+17) at
    // limb 6; the "+1" copy of (2^59+17) was preloaded too, so only n*(2^59+17) is added here.
    const uint32_t n0 = ~T[0], n1 = ~T[1], n2 = ~T[2], n3 = ~T[3], n4 = ~T[4], n5 = ~T[5];
    const uint32_t q0 = S252_P6;
#if !S252_MONT_SHIFT
    const uint32_t q1 = S252_P7;
#endif
    // {n0,n2,n4}*q0 -> limbs 6,8,10 ; carry ripples to limb 15
    asm("mad.lo.cc.u32  %0, %10, %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %10, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %11, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %11, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %12, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %12, %13, %5;\n\t"
        "addc.cc.u32 %6, %6, 0;\n\t"
        "addc.cc.u32 %7, %7, 0;\n\t"
        "addc.cc.u32 %8, %8, 0;\n\t"
        "addc.u32    %9, %9, 0;"
        : "+r"(T[6]), "+r"(T[7]), "+r"(T[8]), "+r"(T[9]), "+r"(T[10]), "+r"(T[11]), "+r"(T[12]), "+r"(T[13]),
          "+r"(T[14]), "+r"(T[15])
        : "r"(n0), "r"(n2), "r"(n4), "r"(q0));
    // {n1,n3,n5}*q0 -> limbs 7,9,11
    asm("mad.lo.cc.u32  %0, %9,  %12, %0;\n\t"
        "madc.hi.cc.u32 %1, %9,  %12, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %12, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %12, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %12, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %12, %5;\n\t"
        "addc.cc.u32 %6, %6, 0;\n\t"
        "addc.cc.u32 %7, %7, 0;\n\t"
        "addc.u32    %8, %8, 0;"
        : "+r"(T[7]), "+r"(T[8]), "+r"(T[9]), "+r"(T[10]), "+r"(T[11]), "+r"(T[12]), "+r"(T[13]), "+r"(T[14]),
          "+r"(T[15])
        : "r"(n1), "r"(n3), "r"(n5), "r"(q0));
#if S252_MONT_SHIFT
    // n*q1 = n*2^27 at limb 7 is a shift: seven funnel shifts and one carry chain on the ALU pipe instead of six wide multiplies
    // on the (4-cycle) multiplier -- they issue in its shadow
    {
        uint32_t x0, x1, x2, x3, x4, x5, x6;
        asm("shl.b32 %0, %7, 27;\n\t"
            "shf.l.clamp.b32 %1, %7,  %8,  27;\n\t"
            "shf.l.clamp.b32 %2, %8,  %9,  27;\n\t"
            "shf.l.clamp.b32 %3, %9,  %10, 27;\n\t"
            "shf.l.clamp.b32 %4, %10, %11, 27;\n\t"
            "shf.l.clamp.b32 %5, %11, %12, 27;\n\t"
            "shr.u32 %6, %12, 5;"
            : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3), "=r"(x4), "=r"(x5), "=r"(x6)
            : "r"(n0), "r"(n1), "r"(n2), "r"(n3), "r"(n4), "r"(n5));
        asm("add.cc.u32  %0, %0, %9;\n\t"
            "addc.cc.u32 %1, %1, %10;\n\t"
            "addc.cc.u32 %2, %2, %11;\n\t"
            "addc.cc.u32 %3, %3, %12;\n\t"
            "addc.cc.u32 %4, %4, %13;\n\t"
            "addc.cc.u32 %5, %5, %14;\n\t"
            "addc.cc.u32 %6, %6, %15;\n\t"
            "addc.cc.u32 %7, %7, 0;\n\t"
            "addc.u32    %8, %8, 0;"
            : "+r"(T[7]), "+r"(T[8]), "+r"(T[9]), "+r"(T[10]), "+r"(T[11]), "+r"(T[12]), "+r"(T[13]), "+r"(T[14]), "+r"(T[15])
            : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(x4), "r"(x5), "r"(x6));
    }
#else
    // {n0,n2,n4}*q1 -> limbs 7,9,11
    asm("mad.lo.cc.u32  %0, %9,  %12, %0;\n\t"
        "madc.hi.cc.u32 %1, %9,  %12, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %12, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %12, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %12, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %12, %5;\n\t"
        "addc.cc.u32 %6, %6, 0;\n\t"
        "addc.cc.u32 %7, %7, 0;\n\t"
        "addc.u32    %8, %8, 0;"
        : "+r"(T[7]), "+r"(T[8]), "+r"(T[9]), "+r"(T[10]), "+r"(T[11]), "+r"(T[12]), "+r"(T[13]), "+r"(T[14]),
          "+r"(T[15])
        : "r"(n0), "r"(n2), "r"(n4), "r"(q1));
    // {n1,n3,n5}*q1 -> limbs 8,10,12
    asm("mad.lo.cc.u32  %0, %8,  %11, %0;\n\t"
        "madc.hi.cc.u32 %1, %8,  %11, %1;\n\t"
        "madc.lo.cc.u32 %2, %9,  %11, %2;\n\t"
        "madc.hi.cc.u32 %3, %9,  %11, %3;\n\t"
        "madc.lo.cc.u32 %4, %10, %11, %4;\n\t"
        "madc.hi.cc.u32 %5, %10, %11, %5;\n\t"
        "addc.cc.u32 %6, %6, 0;\n\t"
        "addc.u32    %7, %7, 0;"
        : "+r"(T[8]), "+r"(T[9]), "+r"(T[10]), "+r"(T[11]), "+r"(T[12]), "+r"(T[13]), "+r"(T[14]), "+r"(T[15])
        : "r"(n1), "r"(n3), "r"(n5), "r"(q1));
#endif
    // Reduction step 2 (64-bit digit): limbs 6,7 are cancelled the same way; (n'+1)*p adds
    // n'*(2^59+17) at limb 12 (the "+1" copy and the carry into limb 8 were preloaded).
    const uint32_t r0 = ~T[6], r1 = ~T[7];
#if S252_MONT_SHIFT
    asm("mad.lo.cc.u32  %0, %4, %5, %0;\n\t"     // r0*q0 -> 12,13
        "madc.hi.cc.u32 %1, %4, %5, %1;\n\t"
        "addc.cc.u32    %2, %2, 0;\n\t"
        "addc.u32       %3, %3, 0;"
        : "+r"(T[12]), "+r"(T[13]), "+r"(T[14]), "+r"(T[15])
        : "r"(r0), "r"(q0));
    asm("mad.lo.cc.u32  %0, %3, %4, %0;\n\t"     // r1*q0 -> 13,14
        "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
        "addc.u32       %2, %2, 0;"
        : "+r"(T[13]), "+r"(T[14]), "+r"(T[15])
        : "r"(r1), "r"(q0));
    {
        uint32_t y0, y1, y2;                      // (r1:r0)*2^27 -> 13,14,15
        asm("shl.b32 %0, %3, 27;\n\t"
            "shf.l.clamp.b32 %1, %3, %4, 27;\n\t"
            "shr.u32 %2, %4, 5;"
            : "=r"(y0), "=r"(y1), "=r"(y2)
            : "r"(r0), "r"(r1));
        asm("add.cc.u32  %0, %0, %3;\n\t"
            "addc.cc.u32 %1, %1, %4;\n\t"
            "addc.u32    %2, %2, %5;"
            : "+r"(T[13]), "+r"(T[14]), "+r"(T[15])
            : "r"(y0), "r"(y1), "r"(y2));
    }
#else
    asm("mad.lo.cc.u32  %0, %4, %6, %0;\n\t"     // r0*q0 -> 12,13
        "madc.hi.cc.u32 %1, %4, %6, %1;\n\t"
        "madc.lo.cc.u32 %2, %5, %7, %2;\n\t"     // r1*q1 -> 14,15
        "madc.hi.u32    %3, %5, %7, %3;"
        : "+r"(T[12]), "+r"(T[13]), "+r"(T[14]), "+r"(T[15])
        : "r"(r0), "r"(r1), "r"(q0), "r"(q1));
    asm("mad.lo.cc.u32  %0, %3, %4, %0;\n\t"     // r1*q0 -> 13,14
        "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
        "addc.u32       %2, %2, 0;"
        : "+r"(T[13]), "+r"(T[14]), "+r"(T[15])
        : "r"(r1), "r"(q0));
    asm("mad.lo.cc.u32  %0, %3, %4, %0;\n\t"     // r0*q1 -> 13,14
        "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
        "addc.u32       %2, %2, 0;"
        : "+r"(T[13]), "+r"(T[14]), "+r"(T[15])
        : "r"(r0), "r"(q1));
#endif
    fe r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.l[i] = T[8 + i];
    return r;
}

// The product rows with no accumulator initialisation at all (generated and checked by tools/gen_fe_mul.py): every first
// touch of a limb is a plain wide product or an immediate addend.  Saves ~20 register moves per multiplication.
#ifndef S252_FE_MUL_GEN
#define S252_FE_MUL_GEN 0
#endif
#include "fe_mul_rows.inc"

// Montgomery product a*b/R mod p, lazily reduced.
// Requires (a/p)*(b/p) <= 31 (e.g. a < 31p with a fully reduced twiddle b); returns a value < 2p.
__device__ __forceinline__ fe fe_mul(const fe& a, const fe& b) {
#if S252_FE_MUL_GEN
    uint32_t E[17], O[16];
    fe_mul_rows(E, O, a.l, b.l);
#else
    // E collects 64-bit products that start on even limbs, O those that start on odd limbs
    // (O[k] has weight 2^(32(k+1))).  The constants preloaded into E are the "+p" and "+1" terms of
    // the two reduction steps (see below); they sit on limbs no product row uses as a carry sink
    // in a way that could overflow.
    uint32_t E[17] = {0, 0, 0, 0, 0, 0, S252_P6 + 1u, S252_P7, 1u, 0, 0, 0, S252_P6, S252_P7, 0, 0, 0};
    uint32_t O[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
        mad_row4(&E[i], &a.l[0], b.l[i]);          // even a * b_i      -> limbs i, i+2, ..
        mad_row4(&O[i], &a.l[1], b.l[i]);          // odd  a * b_i      -> limbs i+1, ..
        mad_row4(&O[i], &a.l[0], b.l[i + 1]);      // even a * b_(i+1)  -> limbs i+1, ..
        if (i < 6) mad_row4(&E[i + 2], &a.l[1], b.l[i + 1]);   // odd a * b_(i+1) -> limbs i+2, ..
        else       mad_row4_nc(&E[i + 2], &a.l[1], b.l[i + 1]); // top row: no carry out of 2^512
    }
#endif
    // T = E + (O << 32): one 16-limb carry chain.
    uint32_t T[16];
    T[0] = E[0];
    asm("add.cc.u32  %0, %15, %30;\n\t"
        "addc.cc.u32 %1, %16, %31;\n\t"
        "addc.cc.u32 %2, %17, %32;\n\t"
        "addc.cc.u32 %3, %18, %33;\n\t"
        "addc.cc.u32 %4, %19, %34;\n\t"
        "addc.cc.u32 %5, %20, %35;\n\t"
        "addc.cc.u32 %6, %21, %36;\n\t"
        "addc.cc.u32 %7, %22, %37;\n\t"
        "addc.cc.u32 %8, %23, %38;\n\t"
        "addc.cc.u32 %9, %24, %39;\n\t"
        "addc.cc.u32 %10, %25, %40;\n\t"
        "addc.cc.u32 %11, %26, %41;\n\t"
        "addc.cc.u32 %12, %27, %42;\n\t"
        "addc.cc.u32 %13, %28, %43;\n\t"
        "addc.u32    %14, %29, %44;"
        : "=r"(T[1]), "=r"(T[2]), "=r"(T[3]), "=r"(T[4]), "=r"(T[5]), "=r"(T[6]), "=r"(T[7]), "=r"(T[8]),
          "=r"(T[9]), "=r"(T[10]), "=r"(T[11]), "=r"(T[12]), "=r"(T[13]), "=r"(T[14]), "=r"(T[15])
        : "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]), "r"(E[8]), "r"(E[9]),
          "r"(E[10]), "r"(E[11]), "r"(E[12]), "r"(E[13]), "r"(E[14]), "r"(E[15]),
          "r"(O[0]), "r"(O[1]), "r"(O[2]), "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]), "r"(O[7]), "r"(O[8]),
          "r"(O[9]), "r"(O[10]), "r"(O[11]), "r"(O[12]), "r"(O[13]), "r"(O[14]));
    return mont_reduce(T);
}

// a + b without reduction (caller tracks the bound; everything must stay below 2^256 ~ 31.99p)
__device__ __forceinline__ fe fe_add_lazy(const fe& a, const fe& b) {
    fe r;
    asm("add.cc.u32  %0, %8,  %16;\n\t"
        "addc.cc.u32 %1, %9,  %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32    %7, %15, %23;"
        : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]), "=r"(r.l[6]), "=r"(r.l[7])
        : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]), "r"(a.l[6]), "r"(a.l[7]),
          "r"(b.l[0]), "r"(b.l[1]), "r"(b.l[2]), "r"(b.l[3]), "r"(b.l[4]), "r"(b.l[5]), "r"(b.l[6]), "r"(b.l[7]));
    return r;
}
// a - b + K*p for b < K*p (result < a + K*p)
template <uint32_t K>
__device__ __forceinline__ fe fe_sub_lazy(const fe& a, const fe& b) {
    static_assert(K >= 1 && K <= 31, "K*p must fit in 256 bits");
    fe t, r;
    asm("sub.cc.u32  %0, %8,  %11;\n\t"
        "subc.cc.u32 %1, 0,   %12;\n\t"
        "subc.cc.u32 %2, 0,   %13;\n\t"
        "subc.cc.u32 %3, 0,   %14;\n\t"
        "subc.cc.u32 %4, 0,   %15;\n\t"
        "subc.cc.u32 %5, 0,   %16;\n\t"
        "subc.cc.u32 %6, %9,  %17;\n\t"
        "subc.u32    %7, %10, %18;"
        : "=r"(t.l[0]), "=r"(t.l[1]), "=r"(t.l[2]), "=r"(t.l[3]), "=r"(t.l[4]), "=r"(t.l[5]), "=r"(t.l[6]), "=r"(t.l[7])
        : "r"(K * S252_P0), "r"(K * S252_P6), "r"(K * S252_P7),
          "r"(b.l[0]), "r"(b.l[1]), "r"(b.l[2]), "r"(b.l[3]), "r"(b.l[4]), "r"(b.l[5]), "r"(b.l[6]), "r"(b.l[7]));
    r = fe_add_lazy(a, t);
    return r;
}

// Bring any value < 2^256 into [0, p).  x = q*2^251 + r with q = x >> 251 <= 31; then
// x - q*p = r - q*(17*2^192 + 1) lies in (-2^202, 2^251), so one conditional +p finishes.
__device__ __forceinline__ fe fe_reduce(const fe& x) {
    const uint32_t q = x.l[7] >> 27;
    uint32_t y[8], borrow;
    asm("sub.cc.u32  %0, %9,  %17;\n\t"
        "subc.cc.u32 %1, %10, 0;\n\t"
        "subc.cc.u32 %2, %11, 0;\n\t"
        "subc.cc.u32 %3, %12, 0;\n\t"
        "subc.cc.u32 %4, %13, 0;\n\t"
        "subc.cc.u32 %5, %14, 0;\n\t"
        "subc.cc.u32 %6, %15, %18;\n\t"
        "subc.cc.u32 %7, %16, 0;\n\t"
        "subc.u32    %8, 0, 0;"
        : "=r"(y[0]), "=r"(y[1]), "=r"(y[2]), "=r"(y[3]), "=r"(y[4]), "=r"(y[5]), "=r"(y[6]), "=r"(y[7]), "=r"(borrow)
        : "r"(x.l[0]), "r"(x.l[1]), "r"(x.l[2]), "r"(x.l[3]), "r"(x.l[4]), "r"(x.l[5]), "r"(x.l[6]),
          "r"(x.l[7] & 0x07ffffffu), "r"(q), "r"(q * S252_P6));
    // borrow = 0xffffffff when the difference went negative
    fe r;
    asm("add.cc.u32  %0, %8,  %16;\n\t"
        "addc.cc.u32 %1, %9,  0;\n\t"
        "addc.cc.u32 %2, %10, 0;\n\t"
        "addc.cc.u32 %3, %11, 0;\n\t"
        "addc.cc.u32 %4, %12, 0;\n\t"
        "addc.cc.u32 %5, %13, 0;\n\t"
        "addc.cc.u32 %6, %14, %17;\n\t"
        "addc.u32    %7, %15, %18;"
        : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]), "=r"(r.l[6]), "=r"(r.l[7])
        : "r"(y[0]), "r"(y[1]), "r"(y[2]), "r"(y[3]), "r"(y[4]), "r"(y[5]), "r"(y[6]), "r"(y[7]),
          "r"(borrow & S252_P0), "r"(borrow & S252_P6), "r"(borrow & S252_P7));
    return r;
}

// fully reduced helpers (used off the hot path: table generation, FRI scalars)
__device__ __forceinline__ fe fe_mul_full(const fe& a, const fe& b) { return fe_reduce(fe_mul(a, b)); }
__device__ __forceinline__ fe fe_add_full(const fe& a, const fe& b) { return fe_reduce(fe_add_lazy(a, b)); }
__device__ __forceinline__ fe fe_sub_full(const fe& a, const fe& b) { return fe_reduce(fe_sub_lazy<1>(a, b)); }
// Montgomery form -> canonical representative in [0, p): a Montgomery product with 1 is the
// reduction alone (16 wide MADs instead of 80) -- used once per element by the leaf hashers.
__device__ __forceinline__ fe fe_from_mont(const fe& a) {
    uint32_t T[16];
#pragma unroll
    for (int i = 0; i < 6; ++i) T[i] = a.l[i];
    // limbs 6.. : a's top two limbs plus the reduction constants (see mont_reduce)
    asm("add.cc.u32  %0, %4, %6;\n\t"
        "addc.cc.u32 %1, %5, %7;\n\t"
        "addc.cc.u32 %2, 1, 0;\n\t"
        "addc.u32    %3, 0, 0;"
        : "=r"(T[6]), "=r"(T[7]), "=r"(T[8]), "=r"(T[9])
        : "r"(a.l[6]), "r"(a.l[7]), "r"(S252_P6 + 1u), "r"(S252_P7));
    T[10] = 0; T[11] = 0; T[12] = S252_P6; T[13] = S252_P7; T[14] = 0; T[15] = 0;
    return fe_reduce(mont_reduce(T));
}
__device__ __forceinline__ fe fe_to_mont(const fe& a) { return fe_reduce(fe_mul(a, fe_r2())); }
__device__ __forceinline__ bool fe_is_zero(const fe& a) {
    return (a.l[0] | a.l[1] | a.l[2] | a.l[3] | a.l[4] | a.l[5] | a.l[6] | a.l[7]) == 0;
}
// a^e, fully reduced, for table generation
__device__ inline fe fe_pow(fe a, uint64_t e) {
    fe r = fe_one();
    while (e) {
        if (e & 1) r = fe_mul_full(r, a);
        a = fe_mul_full(a, a);
        e >>= 1;
    }
    return r;
}

// ---- memory: internal layout is 32 bytes little-endian (two 16-byte halves) ----
__device__ __forceinline__ fe ld_fe(const fe* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 lo = q[0], hi = q[1];
    return fe{{lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w}};
}
__device__ __forceinline__ fe ldg_fe(const fe* p) {   // read-only path
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 lo = __ldg(q), hi = __ldg(q + 1);
    return fe{{lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w}};
}
__device__ __forceinline__ void st_fe(fe* p, const fe& v) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
// "LW" interchange format of the reference: 4 x u64, limbs[0] most significant.  As bytes that is
// u32 words [l1 l0 | l3 l2 | l5 l4 | l7 l6] read from the top: word k of the LW image holds
// internal limb (7 - k) ^ 1.
__device__ __forceinline__ fe ld_lw(const fe* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];   // a = {l6,l7,l4,l5}, b = {l2,l3,l0,l1}
    return fe{{b.z, b.w, b.x, b.y, a.z, a.w, a.x, a.y}};
}
__device__ __forceinline__ void st_lw(fe* p, const fe& v) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(v.l[6], v.l[7], v.l[4], v.l[5]);
    q[1] = make_uint4(v.l[2], v.l[3], v.l[0], v.l[1]);
}

}  // namespace s252
