// stark252_cairo.hpp -- C++ host layer over include/stark252_cairo.h, mirroring the reference's Cairo interface:
//
//   run_program                 src/cairo/runner/run.rs:62-241   (hint-free, builtin-free Cairo 0)
//   build_main_trace            src/cairo/execution_trace.rs:57-87 (+ PublicInputs::from_regs_and_mem, air.rs:183-214)
//   generate_cairo_proof        src/cairo/air.rs:1183-1190       (prove::<CairoAIR>, on the GPU)
//
// Header-only; link with -lstark252_b200.
#pragma once
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

#include "stark252_b200.hpp"
#include "stark252_cairo.h"

namespace stark252 {
namespace cairo {

struct CairoError : std::runtime_error {
    int code;
    CairoError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
inline void check(int rc) { if (rc != S252_OK) throw CairoError(rc, s252_cairo_last_error()); }

// (register_states, memory) in the reference's binary formats (register_states.rs:47-78, cairo_mem.rs:35-61)
struct Execution {
    std::vector<uint8_t> register_states, memory;
    size_t program_size = 0;
};

// program: the `data` words of the compiled JSON, 32 bytes big-endian each
inline Execution run_program(const std::vector<uint8_t>& program_be, uint64_t entry_offset = 0, uint64_t max_steps = 0) {
    s252_cairo_run* r = nullptr;
    check(s252_cairo_vm_run(program_be.data(), program_be.size() / 32, entry_offset, max_steps, &r));
    Execution e;
    e.program_size = program_be.size() / 32;
    e.register_states.resize(s252_cairo_run_trace_len(r));
    e.memory.resize(s252_cairo_run_memory_len(r));
    s252_cairo_run_trace_bytes(r, e.register_states.data());
    s252_cairo_run_memory_bytes(r, e.memory.data());
    s252_cairo_run_destroy(r);
    return e;
}

// TraceTable + PublicInputs of one execution (owned by the library; the table is LW, row-major)
class MainTrace {
  public:
    explicit MainTrace(s252_cairo_trace* h) : h_(h) {}
    MainTrace(MainTrace&& o) noexcept : h_(o.h_) { o.h_ = nullptr; }
    MainTrace(const MainTrace&) = delete;
    ~MainTrace() { if (h_) s252_cairo_trace_destroy(h_); }
    size_t n_rows() const { return s252_cairo_trace_n_rows(h_); }
    size_t n_cols() const { return s252_cairo_trace_n_cols(h_); }
    const FE* table() const { return s252_cairo_trace_table(h_); }
    s252_cairo_public_inputs pub_inputs() const { s252_cairo_public_inputs p; s252_cairo_trace_public_inputs(h_, &p); return p; }
    std::vector<uint8_t> serialize_public_inputs() const {
        std::vector<uint8_t> b(s252_cairo_trace_serialize_public_inputs(h_, nullptr));
        s252_cairo_trace_serialize_public_inputs(h_, b.data());
        return b;
    }
    const s252_cairo_trace* raw() const { return h_; }

  private:
    s252_cairo_trace* h_;
};

inline MainTrace build_main_trace(const Execution& e, const uint64_t* rc_range = nullptr, const uint64_t* output_range = nullptr) {
    s252_cairo_trace* t = nullptr;
    check(s252_cairo_build_main_trace(e.register_states.data(), e.register_states.size(), e.memory.data(), e.memory.size(),
                                      e.program_size, rc_range, output_range, &t));
    return MainTrace(t);
}

// -> StarkProof::serialize() bytes
inline std::vector<uint8_t> generate_cairo_proof(const Context& ctx, const MainTrace& trace, const ProofOptions& options) {
    uint8_t* p = nullptr;
    size_t n = 0;
    ctx.check(s252_cairo_prove(ctx.raw(), trace.raw(), options.blowup_factor, options.fri_number_of_queries, options.coset_offset,
                               options.grinding_factor, &p, &n));
    std::vector<uint8_t> out(p, p + n);
    s252_cairo_proof_free(p);
    return out;
}

// The same proof as ONE collective call over the GPUs of `comm` (every rank calls it with the same trace and options):
// StarkProof::serialize() bytes on rank 0, an empty vector on the other ranks.
inline std::vector<uint8_t> generate_cairo_proof_sharded(const Context& ctx, const Communicator& comm, const MainTrace& trace,
                                                         const ProofOptions& options, size_t pipeline_groups = 0) {
    uint8_t* p = nullptr;
    size_t n = 0;
    ctx.check(s252_cairo_prove_sharded(ctx.raw(), comm.raw(), trace.raw(), options.blowup_factor, options.fri_number_of_queries,
                                       options.coset_offset, options.grinding_factor, pipeline_groups, &p, &n));
    std::vector<uint8_t> out(p, p + n);
    if (p) s252_cairo_proof_free(p);
    return out;
}

}  // namespace cairo
}  // namespace stark252
