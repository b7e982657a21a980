// Exercises the C++ host layer (include/stark252_b200.hpp) the way the reference's own tests read:
// src/starks/prover.rs:838-862 (LDE of the fibonacci trace polys) and a small commit + FRI + grinding.
// Prints values as hex for the python test to compare with the oracle.  Exit code 2 = no GPU.
#include <cstdio>
#include <cstring>
#include <vector>

#include "stark252_b200.hpp"
#include "stark252_cairo.hpp"

using namespace stark252;

static void print_fe(const char* tag, const FE& f) {
    std::printf("%s %016llx%016llx%016llx%016llx\n", tag, (unsigned long long)f.limbs[0], (unsigned long long)f.limbs[1],
                (unsigned long long)f.limbs[2], (unsigned long long)f.limbs[3]);
}
static void print_hex(const char* tag, const uint8_t* b, size_t n) {
    std::printf("%s ", tag);
    for (size_t i = 0; i < n; ++i) std::printf("%02x", b[i]);
    std::printf("\n");
}

int main(int argc, char** argv) {
    if (argc > 1 && std::strcmp(argv[1], "--link-only") == 0) {
        std::printf("len %zu\n", s252_evaluate_offset_fft_len(5, 2, 0));
        return 0;
    }
    if (argc > 1 && std::strcmp(argv[1], "--cairo-front-end") == 0) {
        // host only: the reference's `mul` program (cairo_mem.rs:74-95) through the Cairo machine and build_main_trace
        const uint64_t words[5] = {0x480680017fff8000ULL, 6, 0x400680017fff7fffULL, 6, 0x208b7fff7fff7ffeULL};
        std::vector<uint8_t> prog(5 * 32, 0);
        for (int i = 0; i < 5; ++i)
            for (int k = 0; k < 8; ++k) prog[32 * i + 31 - k] = (uint8_t)(words[i] >> (8 * k));
        auto exe = stark252::cairo::run_program(prog);
        auto trace = stark252::cairo::build_main_trace(exe);
        const auto pi = trace.pub_inputs();
        std::printf("steps %zu rows %zu cols %zu ap_final %llu rc %u %u\n", exe.register_states.size() / 24, trace.n_rows(), trace.n_cols(),
                    (unsigned long long)pi.ap_final, pi.range_check_min, pi.range_check_max);
        if (argc > 2 && std::strcmp(argv[2], "--prove") == 0) {
            Context ctx(0);
            const auto proof = stark252::cairo::generate_cairo_proof(ctx, trace, ProofOptions::default_test_options());
            uint8_t d[32];
            s252_keccak256(proof.data(), proof.size(), d);
            std::printf("proof %zu ", proof.size());
            for (int i = 0; i < 32; ++i) std::printf("%02x", d[i]);
            std::printf("\n");
        }
        return 0;
    }
    try {
        Context ctx(0);
        // trace given on stdin-free deterministic rule: row i, col j -> Montgomery limbs of a small counter pattern
        const size_t n = 16, c = 3;
        TraceTable trace;
        trace.n_cols = c;
        trace.table.resize(n * c);
        for (size_t i = 0; i < n * c; ++i) {
            trace.table[i].limbs[0] = 0x0123456789abcdefULL ^ (i * 0x9e3779b97f4a7c15ULL) >> 8;   // < 2^59: below p
            trace.table[i].limbs[0] &= (1ULL << 59) - 1;
            trace.table[i].limbs[1] = i * 0xbf58476d1ce4e5b9ULL;
            trace.table[i].limbs[2] = ~i;
            trace.table[i].limbs[3] = i + 7;
        }
        ProofOptions opt = ProofOptions::default_test_options();
        DefaultTranscript t;
        Commit commit = interpolate_and_commit(ctx, trace, opt, t);
        print_hex("root", commit.root.data(), 32);
        auto col = commit.lde_column(1);
        print_fe("lde_1_5", col[5]);
        auto proof = commit.get_proof_by_pos(9);
        print_hex("path0", proof->merkle_path[0].data(), 32);
        std::printf("none %d\n", (int)!commit.get_proof_by_pos(n * opt.blowup_factor).has_value());
        Polynomial p0 = commit.trace_poly(2);
        FE h{};
        h = trace.table[0];   // any non-zero element works as a coset offset for the demo
        auto fri = fri_commit_phase(ctx, 4, p0, t, h, n * opt.blowup_factor);
        print_fe("fri_last", fri.first);
        print_hex("fri_root3", fri.second.roots[3].data(), 32);
        auto nonce = generate_nonce_with_grinding(ctx, t.challenge(), 9);
        std::printf("nonce %llu\n", (unsigned long long)*nonce);
        try {
            Polynomial::interpolate_fft(ctx, std::vector<FE>(12));
            std::printf("fft_error 0\n");
        } catch (const FFTError&) {
            std::printf("fft_error 1\n");
        }
    } catch (const Error& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return e.code == S252_ERR_CUDA ? 2 : 1;
    }
    return 0;
}
