// Reduced-radix probe: 9 limbs x 29 bits, 81 plain IMAD.WIDE into 64-bit column sums, then one
// normalisation pass.  Timing only.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
struct f9 { uint32_t l[9]; };
template <int STAGE>
__device__ __forceinline__ f9 mul9(const f9& a, const f9& b) {
    uint64_t c[17];
#pragma unroll
    for (int k = 0; k < 17; ++k) c[k] = 0;
#pragma unroll
    for (int i = 0; i < 9; ++i)
#pragma unroll
        for (int j = 0; j < 9; ++j) c[i + j] += (uint64_t)a.l[i] * b.l[j];
    f9 r;
    if (STAGE == 0) {
#pragma unroll
        for (int i = 0; i < 9; ++i) r.l[i] = ((uint32_t)c[i] ^ (uint32_t)(c[i] >> 32) ^ (uint32_t)c[(i + 8) % 17]) & 0x1fffffffu;
        return r;
    }
    // STAGE 1: carry-normalise all 17 columns to 29-bit limbs (stand-in for reduction + normalisation cost)
    uint64_t carry = 0;
    uint32_t t[18];
#pragma unroll
    for (int k = 0; k < 17; ++k) { uint64_t s = c[k] + carry; t[k] = (uint32_t)s & 0x1fffffffu; carry = s >> 29; }
    t[17] = (uint32_t)carry;
    if (STAGE == 1) {
#pragma unroll
        for (int i = 0; i < 9; ++i) r.l[i] = (t[i] ^ t[i + 9]) & 0x1fffffffu;
        return r;
    }
    // STAGE 2: + 18 more plain MADs (the sparse Montgomery rows) folded in
    uint64_t d[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) d[k] = t[k + 9];
#pragma unroll
    for (int k = 0; k < 6; ++k) { uint32_t m = (~t[k]) & 0x1fffffffu; d[k] += (uint64_t)m * 0x440000u; d[k + 2] += (uint64_t)m * 0x80000u; }
#pragma unroll
    for (int k = 0; k < 3; ++k) { uint32_t m = (~(uint32_t)d[k]) & 0x1fffffffu; d[k + 6] += (uint64_t)m * 0x440000u; if (k + 8 < 9) d[k + 8] += (uint64_t)m * 0x80000u; }
    carry = 0;
#pragma unroll
    for (int k = 0; k < 9; ++k) { uint64_t s = d[k] + carry; r.l[k] = (uint32_t)s & 0x1fffffffu; carry = s >> 29; }
    return r;
}
template <int STAGE>
__global__ void __launch_bounds__(256) kern(int iters, uint32_t* sink) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    f9 a, b, w;
#pragma unroll
    for (int i = 0; i < 9; ++i) { a.l[i] = (tid * 2654435761u + i) & 0x1fffffffu; b.l[i] = (tid * 40503u + i * 77) & 0x1fffffffu; w.l[i] = (sink[tid] + tid * 7919u + i * 0x9e3779b9u) & 0x1fffffffu; }
    for (int it = 0; it < iters; ++it) { a = mul9<STAGE>(a, w); b = mul9<STAGE>(b, w); }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 9; ++i) r ^= a.l[i] ^ b.l[i];
    sink[tid] = r;
}
template <int STAGE> void run(const char* name) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int blocks = sms * 8, iters = 512; uint32_t* sink; cudaMalloc(&sink, (size_t)blocks * 256 * 4); cudaMemset(sink, 0, (size_t)blocks * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int r = 0; r < 4; ++r) { cudaEventRecord(e0); kern<STAGE><<<blocks, 256>>>(iters, sink); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms; }
    double n = (double)blocks * 256 * iters * 2;
    double rate = n / (best * 1e-3);
    printf("%-48s %8.1f G/s   %6.1f cycles per warp-op per SMSP\n", name, rate / 1e9, 148.0 * 4 * 32 * 1.93e9 / rate);
    cudaFree(sink);
}
int main() { run<0>("9x9 plain MADs only"); run<1>("+ 17-column normalisation"); run<2>("+ 17 sparse-reduction MADs + final normalise"); return 0; }
