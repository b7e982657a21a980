// Alternative formulations of the 8x8-limb product phase (timing only; results folded by xor).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../lambdaworks_cairo_prover_b200/csrc/fe.cuh"
using namespace s252;

// V1: current even/odd rows with carry-in/out chains
__device__ __forceinline__ void prod_v1(const fe& a, const fe& b, uint32_t T[16]) {
    uint32_t E[17] = {0}, O[16] = {0};
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
        mad_row4(&E[i], &a.l[0], b.l[i]); mad_row4(&O[i], &a.l[1], b.l[i]); mad_row4(&O[i], &a.l[0], b.l[i + 1]);
        if (i < 6) mad_row4(&E[i + 2], &a.l[1], b.l[i + 1]); else mad_row4_nc(&E[i + 2], &a.l[1], b.l[i + 1]);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) T[i] = E[i] ^ (i ? O[i - 1] : 0);
}
// V2: product scanning, 96-bit column accumulator (mad.lo.cc / madc.hi.cc / addc)
__device__ __forceinline__ void prod_v2(const fe& a, const fe& b, uint32_t T[16]) {
    uint32_t c0 = 0, c1 = 0, c2 = 0;
#pragma unroll
    for (int k = 0; k < 15; ++k) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int j = k - i;
            if (j < 0 || j > 7) continue;
            asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;" : "+r"(c0), "+r"(c1), "+r"(c2) : "r"(a.l[i]), "r"(b.l[j]));
        }
        T[k] = c0; c0 = c1; c1 = c2; c2 = 0;
    }
    T[15] = c0;
}
// V3: plain 64-bit column sums: lo and hi halves of every product summed separately in 64-bit
// accumulators (no carries at all during accumulation), resolved by one carry pass.
__device__ __forceinline__ void prod_v3(const fe& a, const fe& b, uint32_t T[16]) {
    uint64_t lo[15], hi[15];
#pragma unroll
    for (int k = 0; k < 15; ++k) { lo[k] = 0; hi[k] = 0; }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            uint64_t p = (uint64_t)a.l[i] * b.l[j];
            lo[i + j] += (uint32_t)p;
            hi[i + j] += p >> 32;
        }
    uint64_t carry = 0;
#pragma unroll
    for (int k = 0; k < 15; ++k) {
        uint64_t s = lo[k] + (k ? hi[k - 1] : 0) + carry;
        T[k] = (uint32_t)s; carry = s >> 32;
    }
    T[15] = (uint32_t)(hi[14] + carry);
}
// V4: 32-bit mad.lo / mad.hi into 64-bit-free column accumulators using IADD3-friendly 32-bit sums of 16-bit-split... skipped
template <int V>
__global__ void __launch_bounds__(256) kern(int iters, fe* sink) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    fe a = fe_one(), b = fe_r2(), w;
#pragma unroll
    for (int i = 0; i < 8; ++i) w.l[i] = tid * 2654435761u + i * 0x9e3779b9u + sink[tid].l[i];
    a.l[0] ^= tid; b.l[1] ^= tid;
    for (int it = 0; it < iters; ++it) {
        uint32_t T[16], U[16];
        if (V == 1) { prod_v1(a, w, T); prod_v1(b, w, U); }
        if (V == 2) { prod_v2(a, w, T); prod_v2(b, w, U); }
        if (V == 3) { prod_v3(a, w, T); prod_v3(b, w, U); }
#pragma unroll
        for (int i = 0; i < 8; ++i) { a.l[i] = T[i] ^ T[i + 8]; b.l[i] = U[i] ^ U[i + 8]; }
    }
    st_fe(sink + tid, fe_add_lazy(a, b));
}
template <int V> void run(const char* name) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int blocks = sms * 8, iters = 512; fe* sink; cudaMalloc(&sink, (size_t)blocks * 256 * 32); cudaMemset(sink, 0, (size_t)blocks * 256 * 32);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int r = 0; r < 4; ++r) { cudaEventRecord(e0); kern<V><<<blocks, 256>>>(iters, sink); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms; }
    double n = (double)blocks * 256 * iters * 2;
    double rate = n / (best * 1e-3);
    printf("%-44s %8.1f G/s   %6.1f cycles per warp-product per SMSP\n", name, rate / 1e9, 148.0 * 4 * 32 * 1.93e9 / rate);
    cudaFree(sink);
}
int main() { run<1>("V1 even/odd rows, carry chains"); run<2>("V2 product scanning 96-bit acc"); run<3>("V3 split lo/hi 64-bit column sums"); return 0; }
