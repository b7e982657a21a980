// Experimental issue-rate probes for sm_100a integer instructions (not part of the product).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CH 8
#define UN 8
template <int W>
__global__ void __launch_bounds__(256) k(int iters, uint32_t* sink) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t acc[CH]; uint32_t x[CH], y[CH], z[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) { acc[c] = tid * 2654435761u + c; x[c] = tid ^ (0x9e3779b9u * (c + 1)); y[c] = (tid + c) | 1u; z[c] = tid * 7 + c; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < UN; ++u)
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                if (W == 0) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[c]) : "r"(x[c]), "r"(y[c]));
                if (W == 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(y[c]), "r"(z[c]));
                if (W == 2) { // mad.wide + lop3(3 reg)
                    asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[c]) : "r"(y[c]), "r"(z[c]));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(y[c]), "r"(z[c])); }
                if (W == 3) { // mad.wide + add (2 reg)
                    asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[c]) : "r"(y[c]), "r"(z[c]));
                    asm volatile("add.u32 %0, %0, %1;" : "+r"(x[c]) : "r"(y[c])); }
                if (W == 4) { // mad.lo (32-bit IMAD) + lop3
                    asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(z[c]) : "r"(y[c]), "r"(y[(c+1)%CH]));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(y[c]), "r"(y[(c+2)%CH])); }
                if (W == 5) { // lop3 + add (both ALU)
                    asm volatile("add.u32 %0, %0, %1;" : "+r"(z[c]) : "r"(y[c]));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(y[c]), "r"(y[(c+2)%CH])); }
                if (W == 6) { // mad.wide x imm + xor imm
                    asm volatile("mad.wide.u32 %0, %1, 17, %0;" : "+l"(acc[c]) : "r"(y[c]));
                    asm volatile("xor.b32 %0, %0, 0x55;" : "+r"(x[c])); }
                if (W == 7) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(z[c]) : "r"(x[c]), "r"(y[c]));
                if (W == 8) { // 2 mad.wide : 1 lop3
                    asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[c]) : "r"(y[c]), "r"(z[c]));
                    asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[(c+1)%CH]) : "r"(y[c]), "r"(z[c]));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(y[c]), "r"(z[c])); }
                if (W == 9) { // mad.wide + shf
                    asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[c]) : "r"(y[c]), "r"(z[c]));
                    asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(x[c]) : "r"(y[c])); }
                if (W == 10) { // mad.wide carry chain pair (like the field multiply): 2 wide mads with cc
                    asm volatile("mad.lo.cc.u32 %0, %4, %5, %0;\n\tmadc.hi.cc.u32 %1, %4, %5, %1;\n\tmadc.lo.cc.u32 %2, %6, %5, %2;\n\tmadc.hi.u32 %3, %6, %5, %3;"
                                 : "+r"(x[c]), "+r"(x[(c+1)%CH]), "+r"(z[c]), "+r"(z[(c+1)%CH]) : "r"(y[c]), "r"(y[(c+1)%CH]), "r"(y[(c+2)%CH])); }
            }
    }
    uint32_t r = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) r ^= (uint32_t)acc[c] ^ (uint32_t)(acc[c] >> 32) ^ x[c] ^ y[c] ^ z[c];
    sink[tid] = r;
}
template <int W> void run(const char* name, double ops_per_inner) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int blocks = sms * 8, iters = 1024; uint32_t* sink; cudaMalloc(&sink, blocks * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int r = 0; r < 4; ++r) { cudaEventRecord(e0); k<W><<<blocks, 256>>>(iters, sink); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms; }
    double ops = (double)blocks * 256 * iters * UN * CH * ops_per_inner;
    printf("%-34s %8.1f Ginstr/s  (%.1f lanes/clk/SM @1.93GHz)\n", name, ops / (best * 1e-3) / 1e9, ops / (best * 1e-3) / 148 / 1.93e9);
    cudaFree(sink);
}
int main() {
    run<0>("mad.wide", 1); run<1>("lop3 (3 reg)", 1); run<7>("mad.lo (IMAD 32)", 1);
    run<2>("mad.wide + lop3(3reg)", 2); run<3>("mad.wide + add(2reg)", 2); run<4>("mad.lo + lop3", 2);
    run<5>("add + lop3 (both ALU)", 2); run<6>("mad.wide*imm + xor imm", 2); run<8>("2 mad.wide + 1 lop3", 3);
    run<9>("mad.wide + shf", 2); run<10>("mad cc chain (4 instr -> 2 wide)", 2);
    return 0;
}
