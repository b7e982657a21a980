// Register-bank probe for IMAD.WIDE.U32 (not part of the product): the multiplier operands are pinned to known register
// parities by loading them with 128-bit loads (a quad R4k..R4k+3 per load), the accumulators are 64-bit pairs.
//   V0  a even, b odd   (a = q.x, b = q.y)            V1  a, b both = 0 mod 4 (q.x, r.x)
//   V2  a = 0, b = 2 mod 4 (q.x, r.z)                 V3  a, b both odd (q.y, r.y)
//   V4/V5  the field multiply's rows: chains of 4 wide MADs with carry-in/out + 1 addc, a/b parities different / equal
//   V6/V7  as V4/V5 but the shared multiplier b varies per row the way fe_mul walks it (8 rows)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CH 8
template <int V>
__global__ void __launch_bounds__(256) k(int iters, const uint4* __restrict__ src, uint32_t* sink) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint4 q[4], r[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { q[i] = src[(tid * 8 + i) & 1023]; r[i] = src[(tid * 8 + 4 + i) & 1023]; }
    uint64_t acc[CH];
    uint32_t cnt = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) acc[c] = tid * 2654435761u + c;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (V < 4) {
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    const uint4 &qa = q[c & 3], &ra = r[(c + u) & 3];
                    uint32_t a, b;
                    if (V == 0) { a = (c & 4) ? qa.z : qa.x; b = (c & 4) ? q[(c + u) & 3].w : q[(c + u) & 3].y; }
                    if (V == 1) { a = qa.x; b = ra.x; }
                    if (V == 2) { a = qa.x; b = ra.z; }
                    if (V == 3) { a = qa.y; b = ra.y; }
                    asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[c]) : "r"(a), "r"(b));
                }
                // keep the products loop-variant (ptxas would hoist them): bump the multiplicands in place
                if (u == 7) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (V == 0) { asm volatile("add.u32 %0, %0, 2;" : "+r"(q[j].x)); asm volatile("add.u32 %0, %0, 2;" : "+r"(q[j].z)); }
                        if (V == 1 || V == 2) asm volatile("add.u32 %0, %0, 2;" : "+r"(q[j].x));
                        if (V == 3) asm volatile("add.u32 %0, %0, 2;" : "+r"(q[j].y));
                    }
                }
            } else {
                // two rows of 4 chained wide MADs: acc pairs 0..3 and 4..7, four a's, one b per row
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    uint32_t a0, a1, a2, a3, b;
                    if (V == 4 || V == 6) { a0 = q[0].x; a1 = q[1].x; a2 = q[2].x; a3 = q[3].x; b = (V == 4) ? r[g].y : ((u & 1) ? r[(u >> 1) & 3].w : r[(u >> 1) & 3].y); }
                    else                  { a0 = q[0].x; a1 = q[1].x; a2 = q[2].x; a3 = q[3].x; b = (V == 5) ? r[g].x : ((u & 1) ? r[(u >> 1) & 3].z : r[(u >> 1) & 3].x); }
                    uint32_t l0 = (uint32_t)acc[4 * g], h0 = (uint32_t)(acc[4 * g] >> 32), l1 = (uint32_t)acc[4 * g + 1], h1 = (uint32_t)(acc[4 * g + 1] >> 32);
                    uint32_t l2 = (uint32_t)acc[4 * g + 2], h2 = (uint32_t)(acc[4 * g + 2] >> 32), l3 = (uint32_t)acc[4 * g + 3], h3 = (uint32_t)(acc[4 * g + 3] >> 32);
                    asm volatile("mad.lo.cc.u32 %0, %9, %13, %0;\n\tmadc.hi.cc.u32 %1, %9, %13, %1;\n\t"
                                 "madc.lo.cc.u32 %2, %10, %13, %2;\n\tmadc.hi.cc.u32 %3, %10, %13, %3;\n\t"
                                 "madc.lo.cc.u32 %4, %11, %13, %4;\n\tmadc.hi.cc.u32 %5, %11, %13, %5;\n\t"
                                 "madc.lo.cc.u32 %6, %12, %13, %6;\n\tmadc.hi.cc.u32 %7, %12, %13, %7;\n\taddc.u32 %8, %8, 0;"
                                 : "+r"(l0), "+r"(h0), "+r"(l1), "+r"(h1), "+r"(l2), "+r"(h2), "+r"(l3), "+r"(h3), "+r"(cnt)
                                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
                    acc[4 * g] = ((uint64_t)h0 << 32) | l0; acc[4 * g + 1] = ((uint64_t)h1 << 32) | l1;
                    acc[4 * g + 2] = ((uint64_t)h2 << 32) | l2; acc[4 * g + 3] = ((uint64_t)h3 << 32) | l3;
                }
            }
        }
    }
    uint32_t x = cnt;
#pragma unroll
    for (int c = 0; c < CH; ++c) x ^= (uint32_t)acc[c] ^ (uint32_t)(acc[c] >> 32);
#pragma unroll
    for (int i = 0; i < 4; ++i) x ^= q[i].x ^ q[i].y ^ q[i].z ^ q[i].w ^ r[i].x ^ r[i].y ^ r[i].z ^ r[i].w;
    sink[tid] = x;
}
template <int V> void run(const char* name, const uint4* src) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    int blocks = sms * 8, iters = 512; uint32_t* sink; cudaMalloc(&sink, blocks * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int r = 0; r < 4; ++r) { cudaEventRecord(e0); k<V><<<blocks, 256>>>(iters, src, sink); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms; }
    double mads = (double)blocks * 256 * iters * 8 * 8;
    double rate = mads / (best * 1e-3);
    printf("%-60s %8.1f G wide-MAD/s  %5.2f cycles per warp-MAD per SMSP (at %d MHz)\n", name, rate / 1e9, (double)sms * 4 * 32 * clk * 1e3 / rate, clk / 1000);
    cudaFree(sink);
}

// Overlap probe: warps of kind 0 run the carry rows (V4 pattern), warps of kind 1 run 8-limb add chains (IADD3.X) or 3-register
// LOP3s.  M = 0: all rows, 1: all adds, 2: even warps rows / odd warps adds, 3: all lop3, 4: even rows / odd lop3.
template <int M>
__global__ void __launch_bounds__(256) mix(int iters, const uint4* __restrict__ src, uint32_t* sink) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int warp = threadIdx.x >> 5;
    const int kind = (M == 0) ? 0 : (M == 1) ? 1 : (M == 3) ? 2 : (M == 2) ? (warp & 1) : ((warp & 1) ? 2 : 0);
    uint4 q[4], r[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) q[i] = src[(tid * 8 + i) & 1023];
    r[0] = src[(tid * 8 + 4) & 1023]; r[1] = src[(tid * 8 + 5) & 1023];
    uint32_t x[8], y[8], cnt = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) { x[c] = tid * 2654435761u + c; y[c] = tid ^ (0x9e3779b9u * (c + 1)); }
    if (kind == 0) {
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    uint32_t* a = g ? y : x;
                    asm volatile("mad.lo.cc.u32 %0, %9, %13, %0;\n\tmadc.hi.cc.u32 %1, %9, %13, %1;\n\t"
                                 "madc.lo.cc.u32 %2, %10, %13, %2;\n\tmadc.hi.cc.u32 %3, %10, %13, %3;\n\t"
                                 "madc.lo.cc.u32 %4, %11, %13, %4;\n\tmadc.hi.cc.u32 %5, %11, %13, %5;\n\t"
                                 "madc.lo.cc.u32 %6, %12, %13, %6;\n\tmadc.hi.cc.u32 %7, %12, %13, %7;\n\taddc.u32 %8, %8, 0;"
                                 : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]), "+r"(cnt)
                                 : "r"(q[0].x), "r"(q[1].x), "r"(q[2].x), "r"(q[3].x), "r"(r[g].y));
                }
            }
        }
    } else if (kind == 1) {
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int u = 0; u < 8; ++u)
                asm volatile("add.cc.u32 %0, %0, %8;\n\taddc.cc.u32 %1, %1, %9;\n\taddc.cc.u32 %2, %2, %10;\n\taddc.cc.u32 %3, %3, %11;\n\t"
                             "addc.cc.u32 %4, %4, %12;\n\taddc.cc.u32 %5, %5, %13;\n\taddc.cc.u32 %6, %6, %14;\n\taddc.u32 %7, %7, %15;"
                             : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "+r"(x[4]), "+r"(x[5]), "+r"(x[6]), "+r"(x[7])
                             : "r"(y[(u + 0) & 7]), "r"(y[(u + 1) & 7]), "r"(y[(u + 2) & 7]), "r"(y[(u + 3) & 7]), "r"(y[(u + 4) & 7]), "r"(y[(u + 5) & 7]), "r"(y[(u + 6) & 7]), "r"(y[(u + 7) & 7]));
        }
    } else {
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(y[(c + u) & 7]), "r"(y[(c + u + 3) & 7]));
        }
    }
    uint32_t z = cnt;
#pragma unroll
    for (int c = 0; c < 8; ++c) z ^= x[c] ^ y[c];
    sink[tid] = z;
}
template <int M> float runmix(const char* name, const uint4* src) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int blocks = sms * 8, iters = 512; uint32_t* sink; cudaMalloc(&sink, blocks * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int r = 0; r < 4; ++r) { cudaEventRecord(e0); mix<M><<<blocks, 256>>>(iters, src, sink); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms; }
    printf("%-60s %8.3f ms (64 instr per iteration per warp)\n", name, best);
    cudaFree(sink);
    return best;
}
int main() {
    uint4* src; cudaMalloc(&src, 1024 * sizeof(uint4)); cudaMemset(src, 0x5a, 1024 * sizeof(uint4));
    run<0>("V0 plain: a even, b odd", src); run<1>("V1 plain: a, b = 0 mod 4", src); run<2>("V2 plain: a = 0, b = 2 mod 4", src);
    run<3>("V3 plain: a, b odd (1 mod 4)", src);
    run<4>("V4 carry rows: a even, b odd, b fixed per row", src); run<5>("V5 carry rows: a, b even, b fixed per row", src);
    run<6>("V6 carry rows: a even, b odd, b walks", src); run<7>("V7 carry rows: a, b even, b walks", src);
    float t0 = runmix<0>("M0 all warps: carry rows (IMAD.WIDE.X)", src), t1 = runmix<1>("M1 all warps: 8-limb add chains (IADD3.X)", src);
    float t2 = runmix<2>("M2 even warps rows / odd warps adds", src);
    printf("   no overlap would be %.3f ms, full overlap %.3f ms\n", (t0 + t1) / 2, (t0 > t1 ? t0 : t1) / 2);
    float t3 = runmix<3>("M3 all warps: lop3 (3 registers)", src), t4 = runmix<4>("M4 even warps rows / odd warps lop3", src);
    printf("   no overlap would be %.3f ms, full overlap %.3f ms\n", (t0 + t3) / 2, (t0 > t3 ? t0 : t3) / 2);
    return 0;
}
