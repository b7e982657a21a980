// Cost of IMAD.WIDE.U32 variants: plain, carry-out only (+ counter), carry-in/out chains.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CH 8
template <int W>
__global__ void __launch_bounds__(256) k(int iters, uint32_t* sink) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t lo[CH], hi[CH], cnt[CH], a[CH], b[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) { lo[c] = tid * 2654435761u + c; hi[c] = tid ^ (0x9e3779b9u * (c + 1)); cnt[c] = 0; a[c] = (tid + c) | 1u; b[c] = tid * 7 + c + 0x80000000u; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (W == 0) {   // plain mad.wide (no carries): 8 MADs
#pragma unroll
                for (int c = 0; c < CH; ++c) { uint64_t acc = ((uint64_t)hi[c] << 32) | lo[c]; asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(a[c]), "r"(b[(c + u) % CH])); lo[c] = (uint32_t)acc; hi[c] = (uint32_t)(acc >> 32); }
            } else if (W == 1) {   // carry-out only + counter: 8 MADs + 8 addc
#pragma unroll
                for (int c = 0; c < CH; ++c)
                    asm volatile("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;" : "+r"(lo[c]), "+r"(hi[c]), "+r"(cnt[c]) : "r"(a[c]), "r"(b[(c + u) % CH]));
            } else if (W == 2) {   // chains of 4 wide MADs with carry-in/out (the field multiply's rows): 8 MADs + 2 addc
#pragma unroll
                for (int g = 0; g < 2; ++g)
                    asm volatile("mad.lo.cc.u32 %0, %9, %13, %0;\n\tmadc.hi.cc.u32 %1, %9, %13, %1;\n\t"
                                 "madc.lo.cc.u32 %2, %10, %13, %2;\n\tmadc.hi.cc.u32 %3, %10, %13, %3;\n\t"
                                 "madc.lo.cc.u32 %4, %11, %13, %4;\n\tmadc.hi.cc.u32 %5, %11, %13, %5;\n\t"
                                 "madc.lo.cc.u32 %6, %12, %13, %6;\n\tmadc.hi.cc.u32 %7, %12, %13, %7;\n\taddc.u32 %8, %8, 0;"
                                 : "+r"(lo[4 * g]), "+r"(hi[4 * g]), "+r"(lo[4 * g + 1]), "+r"(hi[4 * g + 1]), "+r"(lo[4 * g + 2]), "+r"(hi[4 * g + 2]), "+r"(lo[4 * g + 3]), "+r"(hi[4 * g + 3]), "+r"(cnt[g])
                                 : "r"(a[4 * g]), "r"(a[4 * g + 1]), "r"(a[4 * g + 2]), "r"(a[4 * g + 3]), "r"(b[u]));
            } else if (W == 3) {   // carry-out only, counter via separate predicate-free trick: 8 MADs, no counter (lower bound)
#pragma unroll
                for (int c = 0; c < CH; ++c)
                    asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(lo[c]), "+r"(hi[c]) : "r"(a[c]), "r"(b[(c + u) % CH]));
            }
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) r ^= lo[c] ^ hi[c] ^ cnt[c];
    sink[tid] = r;
}
template <int W> void run(const char* name) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int blocks = sms * 8, iters = 512; uint32_t* sink; cudaMalloc(&sink, blocks * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int r = 0; r < 4; ++r) { cudaEventRecord(e0); k<W><<<blocks, 256>>>(iters, sink); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms; }
    double mads = (double)blocks * 256 * iters * 8 * 8;
    double rate = mads / (best * 1e-3);
    printf("%-52s %8.1f G wide-MAD/s  %5.2f cycles per warp-MAD per SMSP\n", name, rate / 1e9, 148.0 * 4 * 32 * 1.93e9 / rate);
    cudaFree(sink);
}
int main() {
    run<0>("plain mad.wide"); run<3>("wide MAD via mad.lo.cc+madc.hi (no capture)"); run<1>("wide MAD carry-out + addc counter");
    run<2>("chains of 4 wide MADs carry-in/out + 1 addc"); return 0;
}
