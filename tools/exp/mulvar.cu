// Where do the cycles of fe_mul go?  Variants that stop after successive stages (results are wrong
// for STAGE < 3; only the timing matters).  Not part of the product.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../lambdaworks_cairo_prover_b200/csrc/fe.cuh"
using namespace s252;

template <int STAGE>
__device__ __forceinline__ fe mulv(const fe& a, const fe& b) {
    uint32_t E[17] = {0, 0, 0, 0, 0, 0, S252_P6 + 1u, S252_P7, 1u, 0, 0, 0, S252_P6, S252_P7, 0, 0, 0};
    uint32_t O[16] = {0};
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
        mad_row4(&E[i], &a.l[0], b.l[i]);
        mad_row4(&O[i], &a.l[1], b.l[i]);
        mad_row4(&O[i], &a.l[0], b.l[i + 1]);
        if (i < 6) mad_row4(&E[i + 2], &a.l[1], b.l[i + 1]); else mad_row4_nc(&E[i + 2], &a.l[1], b.l[i + 1]);
    }
    fe r;
    if (STAGE == 0) {   // products only: fold E and O with xors (cheap, keeps everything live)
#pragma unroll
        for (int i = 0; i < 8; ++i) r.l[i] = E[i] ^ E[i + 8] ^ O[i] ^ O[i + 7];
        return r;
    }
    uint32_t T[16];
    T[0] = E[0];
    uint32_t c = 0;
    asm("add.cc.u32 %0, %8, %15;\n\taddc.cc.u32 %1, %9, %16;\n\taddc.cc.u32 %2, %10, %17;\n\taddc.cc.u32 %3, %11, %18;\n\t"
        "addc.cc.u32 %4, %12, %19;\n\taddc.cc.u32 %5, %13, %20;\n\taddc.cc.u32 %6, %14, %21;\n\taddc.u32 %7, 0, 0;"
        : "=r"(T[1]), "=r"(T[2]), "=r"(T[3]), "=r"(T[4]), "=r"(T[5]), "=r"(T[6]), "=r"(T[7]), "=r"(c)
        : "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]),
          "r"(O[0]), "r"(O[1]), "r"(O[2]), "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]));
    asm("add.cc.u32 %0, %8, %16;\n\taddc.cc.u32 %1, %9, %17;\n\taddc.cc.u32 %2, %10, %18;\n\taddc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\taddc.cc.u32 %5, %13, %21;\n\taddc.cc.u32 %6, %14, %22;\n\taddc.u32 %7, %15, %23;"
        : "=r"(T[8]), "=r"(T[9]), "=r"(T[10]), "=r"(T[11]), "=r"(T[12]), "=r"(T[13]), "=r"(T[14]), "=r"(T[15])
        : "r"(E[8]), "r"(E[9]), "r"(E[10]), "r"(E[11]), "r"(E[12]), "r"(E[13]), "r"(E[14]), "r"(E[15]),
          "r"(O[7]), "r"(O[8]), "r"(O[9]), "r"(O[10]), "r"(O[11]), "r"(O[12]), "r"(O[13]), "r"(O[14] + c));
    if (STAGE == 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) r.l[i] = T[i] ^ T[i + 8];
        return r;
    }
    return fe_mul(a, b);
}
template <int STAGE>
__global__ void __launch_bounds__(256) kern(int iters, fe* sink) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    fe a = fe_one(), b = fe_r2(), w = fe_r2();
    a.l[0] ^= tid; b.l[1] ^= tid; w.l[2] ^= (tid & 0xffff);
    for (int i = 0; i < iters; ++i) {
        if (STAGE == 3) { a = fe_mul(a, w); b = fe_mul(b, w); }
        else if (STAGE == 4) { fe v = fe_mul(b, w); fe s = fe_add_lazy(a, v); b = fe_reduce(fe_sub_lazy<2>(a, v)); a = fe_reduce(s); }
        else if (STAGE == 5) { fe v = fe_mul(b, w); fe s = fe_add_lazy(a, v); fe d = fe_sub_lazy<2>(a, v); a = s; b = d; a.l[7] &= 0x0fffffff; b.l[7] &= 0x0fffffff; }
        else { a = mulv<STAGE>(a, w); b = mulv<STAGE>(b, w); }
    }
    st_fe(sink + tid, fe_add_lazy(a, b));
}
template <int STAGE> void run(const char* name, double per_iter) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int blocks = sms * 8, iters = 512; fe* sink; cudaMalloc(&sink, (size_t)blocks * 256 * 32);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int r = 0; r < 4; ++r) { cudaEventRecord(e0); kern<STAGE><<<blocks, 256>>>(iters, sink); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms; }
    double n = (double)blocks * 256 * iters * per_iter;
    double rate = n / (best * 1e-3);
    printf("%-44s %8.1f G/s   %6.1f cycles per warp-op per SMSP\n", name, rate / 1e9, 148.0 * 4 * 32 * 1.93e9 / rate);
    cudaFree(sink);
}
int main() {
    run<0>("64 product MADs + captures only", 2);
    run<1>("+ E/O merge", 2);
    run<3>("full fe_mul", 2);
    run<5>("butterfly (mul + add + sub)", 1);
    run<4>("butterfly + 2 full reductions", 1);
    return 0;
}
