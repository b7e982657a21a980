// microbench.cuh -- integer-pipe peak probes and element-wise test kernels.
//
// MEASURED_PEAKS.json carries HBM and tensor numbers only; this path is bound by the integer
// pipes (IMAD.WIDE for the field multiply, LOP3/SHF for Keccak), so their issue peaks are measured
// on the box with independent register-resident chains and reported beside every fraction.
#pragma once
#include "fe.cuh"
#include "keccak.cuh"

namespace s252 {

constexpr int INT_BENCH_CHAINS = 8;
constexpr int INT_BENCH_UNROLL = 8;
constexpr int INT_BENCH_OPS_PER_ITER = INT_BENCH_CHAINS * INT_BENCH_UNROLL;

// which: 0 = IMAD.WIDE.U32 as pure products, 1 = lop3, 2 = shf (funnel shift), 3 = add.cc/addc pair counted per instruction,
// 4 = pure wide products and LOP3 interleaved 1:1 (both counted), 5 = the field multiply's rows (chains of four wide MADs with
// carry-in/out + one addc; wide MADs counted), 6 = IMAD (32-bit mad.lo), 7 = wide products whose two multiplicands sit in
// registers of the same parity, each followed by one xor (both counted, like 4).
//
// Every multiplicand is DATA-DEPENDENT (it comes out of the previous product of its chain).  The round-1 probe multiplied
// loop-invariant registers: ptxas hoisted the 64 products out of the loop and left IADD3 + IADD3.X pairs in it (visible in the
// SASS: IMAD.WIDE .., RZ before the loop, adds inside), so its "IMAD.WIDE peak" of 64 lanes/clk/SM was the rate of two adds.
__global__ void __launch_bounds__(256) int_pipe_bench(int which, int iters, uint32_t* sink) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t lo[INT_BENCH_CHAINS], hi[INT_BENCH_CHAINS];
    uint32_t x[INT_BENCH_CHAINS], y[INT_BENCH_CHAINS];
    uint32_t cnt = 0;
#pragma unroll
    for (int c = 0; c < INT_BENCH_CHAINS; ++c) {
        lo[c] = (tid * 2654435761u + c) | 1u;
        hi[c] = (tid ^ (0x9e3779b9u * (c + 1))) | 0x80000001u;
        x[c] = tid ^ (0x9e3779b9u * (c + 1));
        y[c] = (tid + c) | 1u;
    }
    if (which == 0) {
        // (hi:lo) = lo * hi: the product feeds both multiplicands of the next one; a 64-bit pair is (even, odd) registers
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int u = 0; u < INT_BENCH_UNROLL; ++u)
#pragma unroll
                for (int c = 0; c < INT_BENCH_CHAINS; ++c)
                    asm volatile("{\n\t.reg .u64 p;\n\tmul.wide.u32 p, %0, %1;\n\tmov.b64 {%0, %1}, p;\n\t}" : "+r"(lo[c]), "+r"(hi[c]));
        }
    } else if (which == 1) {
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int u = 0; u < INT_BENCH_UNROLL; ++u)
#pragma unroll
                for (int c = 0; c < INT_BENCH_CHAINS; ++c)
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(y[c]), "r"(y[(c + 1) % INT_BENCH_CHAINS]));
        }
    } else if (which == 2) {
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int u = 0; u < INT_BENCH_UNROLL; ++u)
#pragma unroll
                for (int c = 0; c < INT_BENCH_CHAINS; ++c)
                    asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(x[c]) : "r"(y[c]));
        }
    } else if (which == 3) {
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int u = 0; u < INT_BENCH_UNROLL / 2; ++u)
#pragma unroll
                for (int c = 0; c < INT_BENCH_CHAINS; ++c)
                    asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;"
                                 : "+r"(x[c]), "+r"(y[c])
                                 : "r"(y[(c + 1) % INT_BENCH_CHAINS]), "r"(x[(c + 3) % INT_BENCH_CHAINS]));
        }
    } else if (which == 4) {
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int u = 0; u < INT_BENCH_UNROLL / 2; ++u)
#pragma unroll
                for (int c = 0; c < INT_BENCH_CHAINS; ++c) {
                    asm volatile("{\n\t.reg .u64 p;\n\tmul.wide.u32 p, %0, %1;\n\tmov.b64 {%0, %1}, p;\n\t}" : "+r"(lo[c]), "+r"(hi[c]));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(y[c]), "r"(y[(c + 2) % INT_BENCH_CHAINS]));
                }
        }
    } else if (which == 5) {
        // two rows per step: accumulator pairs 0..3 and 4..7, the multiplicands are last step's low words
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int u = 0; u < INT_BENCH_UNROLL; ++u) {
                uint32_t m[INT_BENCH_CHAINS];
#pragma unroll
                for (int c = 0; c < INT_BENCH_CHAINS; ++c) m[c] = lo[(c + 1 + u) % INT_BENCH_CHAINS];
#pragma unroll
                for (int g = 0; g < 2; ++g)
                    asm volatile("mad.lo.cc.u32 %0, %9, %13, %0;\n\tmadc.hi.cc.u32 %1, %9, %13, %1;\n\t"
                                 "madc.lo.cc.u32 %2, %10, %13, %2;\n\tmadc.hi.cc.u32 %3, %10, %13, %3;\n\t"
                                 "madc.lo.cc.u32 %4, %11, %13, %4;\n\tmadc.hi.cc.u32 %5, %11, %13, %5;\n\t"
                                 "madc.lo.cc.u32 %6, %12, %13, %6;\n\tmadc.hi.cc.u32 %7, %12, %13, %7;\n\taddc.u32 %8, %8, 0;"
                                 : "+r"(lo[4 * g]), "+r"(hi[4 * g]), "+r"(lo[4 * g + 1]), "+r"(hi[4 * g + 1]), "+r"(lo[4 * g + 2]),
                                   "+r"(hi[4 * g + 2]), "+r"(lo[4 * g + 3]), "+r"(hi[4 * g + 3]), "+r"(cnt)
                                 : "r"(m[4 * g]), "r"(m[4 * g + 1]), "r"(m[4 * g + 2]), "r"(m[4 * g + 3]), "r"(y[(g + u) % INT_BENCH_CHAINS]));
            }
        }
    } else if (which == 6) {
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int u = 0; u < INT_BENCH_UNROLL; ++u)
#pragma unroll
                for (int c = 0; c < INT_BENCH_CHAINS; ++c)
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(y[c]), "r"(y[(c + 1) % INT_BENCH_CHAINS]));
        }
    } else {
        // lo[c] * lo[c+1]: both multiplicands are the low (even) halves of 64-bit pairs; the high half is folded into hi[c] with
        // one xor so that the full product stays live (compare with which == 4: the same instruction mix, opposite parities)
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int u = 0; u < INT_BENCH_UNROLL / 2; ++u) {
                uint32_t m[INT_BENCH_CHAINS];
#pragma unroll
                for (int c = 0; c < INT_BENCH_CHAINS; ++c) m[c] = lo[(c + 1) % INT_BENCH_CHAINS];
#pragma unroll
                for (int c = 0; c < INT_BENCH_CHAINS; ++c)
                    asm volatile("{\n\t.reg .u64 p;\n\t.reg .u32 h;\n\tmul.wide.u32 p, %0, %2;\n\tmov.b64 {%0, h}, p;\n\txor.b32 %1, %1, h;\n\t}"
                                 : "+r"(lo[c]), "+r"(hi[c]) : "r"(m[c]));
            }
        }
    }
    uint32_t r = cnt;
#pragma unroll
    for (int c = 0; c < INT_BENCH_CHAINS; ++c) r ^= lo[c] ^ hi[c] ^ x[c] ^ y[c];
    sink[tid] = r;
}

// two independent dependent-chains of fe_mul per thread
__global__ void __launch_bounds__(256) fe_mul_bench(int iters, fe* sink) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    fe a = fe_one(), b = fe_r2(), w = fe_r2();
    a.l[0] ^= tid;
    b.l[1] ^= tid;
    w.l[2] ^= (tid & 0xffff);
    for (int i = 0; i < iters; ++i) {
        a = fe_mul(a, w);
        b = fe_mul(b, w);
    }
    st_fe(sink + tid, fe_reduce(fe_add_lazy(a, b)));
}

__global__ void __launch_bounds__(128) keccak_bench(int iters, uint64_t* sink) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t st[25];
#pragma unroll
    for (int k = 0; k < 25; ++k) st[k] = (uint64_t)tid * 0x9e3779b97f4a7c15ULL + k;
    for (int i = 0; i < iters; ++i) keccak_f1600(st);
    uint64_t r = 0;
#pragma unroll
    for (int k = 0; k < 25; ++k) r ^= st[k];
    sink[tid] = r;
}

// element-wise field ops on LW buffers (parity tests of the device arithmetic)
__global__ void fe_binop_kernel(int op, const fe* __restrict__ a, const fe* __restrict__ b, fe* __restrict__ out,
                                unsigned long long n) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const fe x = ld_lw(a + i), y = ld_lw(b + i);
    fe r;
    if (op == 0) r = fe_mul_full(x, y);
    else if (op == 1) r = fe_add_full(x, y);
    else if (op == 2) r = fe_sub_full(x, y);
    else {
        // x^(p-2), p-2 = 0x0800000000000010 ffff.. (192 ones)
        r = fe_one();
        fe base = x;
        for (int bit = 0; bit < 252; ++bit) {
            const bool set = bit < 192 || bit == 196 || bit == 251;
            if (set) r = fe_mul_full(r, base);
            base = fe_mul_full(base, base);
        }
    }
    st_lw(out + i, r);
}

// Keccak256 of arbitrary-length byte strings (one thread per message; test helper, byte-wise)
__global__ void keccak_bytes_kernel(const uint8_t* __restrict__ msgs, unsigned long long msg_len, unsigned long long n,
                                    uint8_t* __restrict__ digests) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t st[25];
#pragma unroll
    for (int k = 0; k < 25; ++k) st[k] = 0;
    const uint8_t* m = msgs + i * msg_len;
    unsigned long long off = 0;
    uint64_t block[17];
    while (true) {
        const unsigned long long left = msg_len - off;
        const bool last = left < 136;
#pragma unroll
        for (int k = 0; k < 17; ++k) block[k] = 0;
        const unsigned take = last ? (unsigned)left : 136u;
        for (unsigned bpos = 0; bpos < take; ++bpos) {
            const uint64_t byte = m[off + bpos];
            const unsigned lane = bpos >> 3;
#pragma unroll
            for (int k = 0; k < 17; ++k)
                if (k == (int)lane) block[k] |= byte << (8 * (bpos & 7));
        }
        if (last) {
            const unsigned lane = take >> 3;
#pragma unroll
            for (int k = 0; k < 17; ++k)
                if (k == (int)lane) block[k] ^= (uint64_t)0x01 << (8 * (take & 7));
            block[16] ^= 0x8000000000000000ULL;
        }
#pragma unroll
        for (int k = 0; k < 17; ++k) st[k] ^= block[k];
        keccak_f1600(st);
        if (last) break;
        off += 136;
    }
    uint8_t* d = digests + 32 * i;
    for (int k = 0; k < 32; ++k) d[k] = (uint8_t)(st[k >> 3] >> (8 * (k & 7)));
}

}  // namespace s252
