// Cost of the wide multiply-add flavours with DATA-DEPENDENT multiplicands (ptxas hoists or strength-reduces loop-invariant
// products, which made earlier probes measure adds).  Every multiplicand x comes from last iteration's accumulators.
//   W0 plain mad.wide (64-bit addend, no flags)                 W1 wide MAD with carry-out captured by an addc counter
//   W2 chains of 2 wide MADs carry-in/out + addc                 W3 chains of 4 (the field multiply's rows) + addc
//   W4 chains of 8 + addc                                        W5 mul.wide (no addend) + add.cc/addc          (64-bit acc)
//   W6 mul.wide + add.cc/addc.cc/addc (96-bit acc)               W7 mul.wide only, xor-folded (pure product rate)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int W>
__global__ void __launch_bounds__(256) k(int iters, const uint4* __restrict__ src, uint32_t* sink) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t lo[8], hi[8], b[8], top[8], cnt = 0;
    {
        uint4 t0 = src[(tid * 4) & 1023], t1 = src[(tid * 4 + 1) & 1023];
        b[0] = t0.x | 1; b[1] = t0.y; b[2] = t0.z; b[3] = t0.w; b[4] = t1.x; b[5] = t1.y; b[6] = t1.z; b[7] = t1.w | 1;
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) { lo[c] = tid * 2654435761u + c; hi[c] = tid ^ (0x9e3779b9u * (c + 1)); top[c] = 0; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            uint32_t x[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) x[c] = lo[(c + 1 + u) & 7];            // data-dependent multiplicands
            if (W == 0) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    uint64_t acc = ((uint64_t)hi[c] << 32) | lo[c];
                    asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(x[c]), "r"(b[(c + u) & 7]));
                    lo[c] = (uint32_t)acc; hi[c] = (uint32_t)(acc >> 32);
                }
            } else if (W == 1) {
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    asm volatile("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;"
                                 : "+r"(lo[c]), "+r"(hi[c]), "+r"(top[c]) : "r"(x[c]), "r"(b[(c + u) & 7]));
            } else if (W == 2) {
#pragma unroll
                for (int g = 0; g < 4; ++g)
                    asm volatile("mad.lo.cc.u32 %0, %5, %7, %0;\n\tmadc.hi.cc.u32 %1, %5, %7, %1;\n\t"
                                 "madc.lo.cc.u32 %2, %6, %7, %2;\n\tmadc.hi.cc.u32 %3, %6, %7, %3;\n\taddc.u32 %4, %4, 0;"
                                 : "+r"(lo[2 * g]), "+r"(hi[2 * g]), "+r"(lo[2 * g + 1]), "+r"(hi[2 * g + 1]), "+r"(top[g])
                                 : "r"(x[2 * g]), "r"(x[2 * g + 1]), "r"(b[(g + u) & 7]));
            } else if (W == 3) {
#pragma unroll
                for (int g = 0; g < 2; ++g)
                    asm volatile("mad.lo.cc.u32 %0, %9, %13, %0;\n\tmadc.hi.cc.u32 %1, %9, %13, %1;\n\t"
                                 "madc.lo.cc.u32 %2, %10, %13, %2;\n\tmadc.hi.cc.u32 %3, %10, %13, %3;\n\t"
                                 "madc.lo.cc.u32 %4, %11, %13, %4;\n\tmadc.hi.cc.u32 %5, %11, %13, %5;\n\t"
                                 "madc.lo.cc.u32 %6, %12, %13, %6;\n\tmadc.hi.cc.u32 %7, %12, %13, %7;\n\taddc.u32 %8, %8, 0;"
                                 : "+r"(lo[4 * g]), "+r"(hi[4 * g]), "+r"(lo[4 * g + 1]), "+r"(hi[4 * g + 1]), "+r"(lo[4 * g + 2]), "+r"(hi[4 * g + 2]),
                                   "+r"(lo[4 * g + 3]), "+r"(hi[4 * g + 3]), "+r"(top[g])
                                 : "r"(x[4 * g]), "r"(x[4 * g + 1]), "r"(x[4 * g + 2]), "r"(x[4 * g + 3]), "r"(b[(g + u) & 7]));
            } else if (W == 4) {
                asm volatile("mad.lo.cc.u32 %0, %17, %25, %0;\n\tmadc.hi.cc.u32 %1, %17, %25, %1;\n\t"
                             "madc.lo.cc.u32 %2, %18, %25, %2;\n\tmadc.hi.cc.u32 %3, %18, %25, %3;\n\t"
                             "madc.lo.cc.u32 %4, %19, %25, %4;\n\tmadc.hi.cc.u32 %5, %19, %25, %5;\n\t"
                             "madc.lo.cc.u32 %6, %20, %25, %6;\n\tmadc.hi.cc.u32 %7, %20, %25, %7;\n\t"
                             "madc.lo.cc.u32 %8, %21, %25, %8;\n\tmadc.hi.cc.u32 %9, %21, %25, %9;\n\t"
                             "madc.lo.cc.u32 %10, %22, %25, %10;\n\tmadc.hi.cc.u32 %11, %22, %25, %11;\n\t"
                             "madc.lo.cc.u32 %12, %23, %25, %12;\n\tmadc.hi.cc.u32 %13, %23, %25, %13;\n\t"
                             "madc.lo.cc.u32 %14, %24, %25, %14;\n\tmadc.hi.cc.u32 %15, %24, %25, %15;\n\taddc.u32 %16, %16, 0;"
                             : "+r"(lo[0]), "+r"(hi[0]), "+r"(lo[1]), "+r"(hi[1]), "+r"(lo[2]), "+r"(hi[2]), "+r"(lo[3]), "+r"(hi[3]),
                               "+r"(lo[4]), "+r"(hi[4]), "+r"(lo[5]), "+r"(hi[5]), "+r"(lo[6]), "+r"(hi[6]), "+r"(lo[7]), "+r"(hi[7]), "+r"(cnt)
                             : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7]), "r"(b[u]));
            } else if (W == 5) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    uint64_t p;
                    asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(x[c]), "r"(b[(c + u) & 7]));
                    asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(lo[c]), "+r"(hi[c]) : "r"((uint32_t)p), "r"((uint32_t)(p >> 32)));
                }
            } else if (W == 6) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    uint64_t p;
                    asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(x[c]), "r"(b[(c + u) & 7]));
                    asm volatile("add.cc.u32 %0, %0, %3;\n\taddc.cc.u32 %1, %1, %4;\n\taddc.u32 %2, %2, 0;"
                                 : "+r"(lo[c]), "+r"(hi[c]), "+r"(top[c]) : "r"((uint32_t)p), "r"((uint32_t)(p >> 32)));
                }
            } else {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    uint64_t p;
                    asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(x[c]), "r"(b[(c + u) & 7]));
                    lo[c] = (uint32_t)p; hi[c] ^= (uint32_t)(p >> 32);
                }
            }
        }
    }
    uint32_t z = cnt;
#pragma unroll
    for (int c = 0; c < 8; ++c) z ^= lo[c] ^ hi[c] ^ top[c];
    sink[tid] = z;
}
template <int W> void run(const char* name, const uint4* src) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    int blocks = sms * 8, iters = 512; uint32_t* sink; cudaMalloc(&sink, blocks * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int r = 0; r < 4; ++r) { cudaEventRecord(e0); k<W><<<blocks, 256>>>(iters, src, sink); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms; }
    double mads = (double)blocks * 256 * iters * 8 * 8;
    double rate = mads / (best * 1e-3);
    printf("%-64s %8.1f G MAD/s  %5.2f cycles per warp-MAD per SMSP\n", name, rate / 1e9, (double)sms * 4 * 32 * clk * 1e3 / rate);
    cudaFree(sink);
}
int main() {
    uint4* src; cudaMalloc(&src, 1024 * sizeof(uint4)); cudaMemset(src, 0x5a, 1024 * sizeof(uint4));
    run<0>("W0 plain mad.wide, 64-bit addend", src); run<1>("W1 wide MAD carry-out + addc", src);
    run<2>("W2 chains of 2 carry-in/out + addc", src); run<3>("W3 chains of 4 carry-in/out + addc", src);
    run<4>("W4 chain of 8 carry-in/out + addc", src); run<5>("W5 mul.wide + add.cc/addc (64-bit acc)", src);
    run<6>("W6 mul.wide + add.cc/addc.cc/addc (96-bit acc)", src); run<7>("W7 mul.wide only", src);
    return 0;
}
