// ONE Cairo proof on several GPUs from a C++ host with ONE THREAD PER GPU (the way a Rust integrator would drive it): every thread
// owns a context and its end of the communicator and calls generate_cairo_proof_sharded (s252_cairo_prove_sharded: NCCL inside the
// library); rank 0 compares the bytes with the single-GPU prover's.  The trace is the reference's `mul` program (cairo_mem.rs:74-95),
// pinned once and shared by all threads.  usage: sharded_prove_demo [n_gpus]   Exit code 2 = not enough GPUs.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

#include "stark252_b200.hpp"
#include "stark252_cairo.hpp"

using namespace stark252;

int main(int argc, char** argv) {
    const int world = argc > 1 ? std::atoi(argv[1]) : 1;
    const uint64_t words[5] = {0x480680017fff8000ULL, 6, 0x400680017fff7fffULL, 6, 0x208b7fff7fff7ffeULL};
    std::vector<uint8_t> prog(5 * 32, 0);
    for (int i = 0; i < 5; ++i)
        for (int k = 0; k < 8; ++k) prog[32 * i + 31 - k] = (uint8_t)(words[i] >> (8 * k));
    auto trace = cairo::build_main_trace(cairo::run_program(prog));
    const ProofOptions opts = ProofOptions::default_test_options();
    std::vector<uint8_t> want;
    try {
        Context ctx0(0);
        s252_cairo_trace_pin(trace.raw());
        want = cairo::generate_cairo_proof(ctx0, trace, opts);
    } catch (const Error& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 2;
    }
    const CommId id = comm_unique_id();
    std::vector<std::string> errors(world);
    std::vector<std::vector<uint8_t>> proofs(world);
    std::vector<std::thread> threads;
    for (int r = 0; r < world; ++r)
        threads.emplace_back([&, r] {
            try {
                Context ctx(r);
                Communicator comm(ctx, id, r, world);
                for (int rep = 0; rep < 2; ++rep) proofs[r] = cairo::generate_cairo_proof_sharded(ctx, comm, trace, opts, rep + 1);
            } catch (const std::exception& e) {
                errors[r] = e.what();
            }
        });
    for (auto& t : threads) t.join();
    for (int r = 0; r < world; ++r)
        if (!errors[r].empty()) { std::fprintf(stderr, "rank %d: %s\n", r, errors[r].c_str()); return errors[r].find("no usable CUDA device") != std::string::npos ? 2 : 1; }
    if (proofs[0] != want) { std::fprintf(stderr, "sharded proof differs from the single-GPU proof (%zu vs %zu bytes)\n", proofs[0].size(), want.size()); return 1; }
    for (int r = 1; r < world; ++r)
        if (!proofs[r].empty()) { std::fprintf(stderr, "rank %d returned bytes\n", r); return 1; }
    std::printf("SHARDED_PROVE_DEMO_OK ranks %d bytes %zu\n", world, want.size());
    return 0;
}
