// stark252_b200.hpp -- C++ host layer over the C ABI (stark252_b200.h), mirroring the names and
// argument meaning of the reference's Rust interface for the LDE + commitment path:
//
//   Polynomial::interpolate_fft / evaluate_offset_fft / interpolate_offset_fft   (lambdaworks-math FFTPoly;
//        call sites src/starks/trace.rs:107, prover.rs:117, fri_commitment.rs:36, evaluation_table.rs:32)
//   evaluate_polynomial_on_lde_domain                                             (src/starks/prover.rs:106-123)
//   BatchedMerkleTree::build, root, get_proof_by_pos                              (src/starks/config.rs:19-20)
//   interpolate_and_commit                                                        (src/starks/prover.rs:126-159)
//   fri_commit_phase, FriLayer                                                    (src/starks/fri/mod.rs:20-72)
//   generate_nonce_with_grinding                                                  (src/starks/grinding.rs:40-48)
//   DefaultTranscript, transcript_to_field, transcript_to_usize                   (src/starks/transcript.rs)
//   ProofOptions                                                                  (src/starks/proof/options.rs:21-26)
//
// Rust's Result<_, FFTError> / Option become exceptions (FFTError, Error) / std::optional.
// Header-only; link with -lstark252_b200.
#pragma once
#include <array>
#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "stark252_b200.h"

namespace stark252 {

using FE = s252_fe;                              // FieldElement<Stark252PrimeField>, in-memory layout of the reference
using Commitment = std::array<uint8_t, 32>;      // src/starks/config.rs:16-17

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
struct FFTError : Error { using Error::Error; };

struct ProofOptions {                            // src/starks/proof/options.rs:21-26
    uint8_t blowup_factor;
    size_t fri_number_of_queries;
    uint64_t coset_offset;
    uint8_t grinding_factor;
    static ProofOptions default_test_options() { return {4, 3, 3, 1}; }   // options.rs:144-151
};

class Context {
  public:
    explicit Context(int device = 0) {
        int rc = s252_ctx_create(device, &ctx_);
        if (rc != S252_OK) throw Error(rc, "s252_ctx_create: no usable CUDA device (there is no CPU fallback)");
    }
    ~Context() { s252_ctx_destroy(ctx_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    s252_ctx* raw() const { return ctx_; }
    template <class E = Error>
    void check(int rc) const { if (rc != S252_OK) throw E(rc, s252_last_error(ctx_)); }

  private:
    s252_ctx* ctx_ = nullptr;
};

// One rank's end of a communicator over the GPUs of one box (one thread or process per GPU; NCCL is called from the library).
// Rank 0 makes the id, every rank constructs its Communicator with it.
using CommId = std::array<uint8_t, S252_COMM_ID_BYTES>;
inline CommId comm_unique_id() {
    CommId id{};
    if (s252_comm_unique_id(id.data()) != S252_OK) throw Error(S252_ERR_CUDA, "NCCL is not available (libnccl.so.2; S252_NCCL_LIB overrides the name)");
    return id;
}
class Communicator {
  public:
    Communicator(const Context& ctx, const CommId& id, int rank, int world) { ctx.check(s252_comm_create(ctx.raw(), id.data(), rank, world, &c_)); }
    ~Communicator() { s252_comm_destroy(c_); }
    Communicator(const Communicator&) = delete;
    Communicator& operator=(const Communicator&) = delete;
    s252_comm* raw() const { return c_; }
    int rank() const { return s252_comm_rank(c_); }
    int world() const { return s252_comm_world(c_); }

  private:
    s252_comm* c_ = nullptr;
};

class DefaultTranscript {                        // lambdaworks_crypto::fiat_shamir::default_transcript
  public:
    DefaultTranscript() : t_(s252_transcript_new()) {}
    ~DefaultTranscript() { s252_transcript_free(t_); }
    DefaultTranscript(const DefaultTranscript&) = delete;
    DefaultTranscript& operator=(const DefaultTranscript&) = delete;
    void append(const uint8_t* data, size_t len) { s252_transcript_append(t_, data, len); }
    void append(const Commitment& c) { append(c.data(), c.size()); }
    Commitment challenge() { Commitment c; s252_transcript_challenge(t_, c.data()); return c; }
    s252_transcript* raw() const { return t_; }

  private:
    s252_transcript* t_;
};
inline FE transcript_to_field(DefaultTranscript& t) { FE f; s252_transcript_to_field(t.raw(), &f); return f; }
inline size_t transcript_to_usize(DefaultTranscript& t) { return (size_t)s252_transcript_to_usize(t.raw()); }

class Polynomial {                               // coefficients low -> high, trailing zeros trimmed (Polynomial::new)
  public:
    Polynomial() = default;
    explicit Polynomial(std::vector<FE> c) : coeffs_(std::move(c)) {
        while (!coeffs_.empty() && !(coeffs_.back().limbs[0] | coeffs_.back().limbs[1] | coeffs_.back().limbs[2] | coeffs_.back().limbs[3]))
            coeffs_.pop_back();
    }
    const std::vector<FE>& coefficients() const { return coeffs_; }
    size_t coeff_len() const { return coeffs_.size(); }

    static Polynomial interpolate_fft(const Context& ctx, const std::vector<FE>& fft_evals) {
        std::vector<FE> out(fft_evals.size());
        ctx.check<FFTError>(s252_interpolate_fft(ctx.raw(), fft_evals.data(), fft_evals.size(), out.data(), S252_HOST));
        return Polynomial(std::move(out));
    }
    static Polynomial interpolate_offset_fft(const Context& ctx, const std::vector<FE>& fft_evals, const FE& offset) {
        std::vector<FE> out(fft_evals.size());
        ctx.check<FFTError>(s252_interpolate_offset_fft(ctx.raw(), fft_evals.data(), fft_evals.size(), &offset, out.data(), S252_HOST));
        return Polynomial(std::move(out));
    }
    // domain_size: std::nullopt == Rust's None
    std::vector<FE> evaluate_offset_fft(const Context& ctx, size_t blowup_factor, std::optional<size_t> domain_size, const FE& offset) const {
        const size_t ds = domain_size.value_or(0);
        std::vector<FE> out(s252_evaluate_offset_fft_len(coeffs_.size(), blowup_factor, ds));
        ctx.check<FFTError>(s252_evaluate_offset_fft(ctx.raw(), coeffs_.data(), coeffs_.size(), blowup_factor, ds, &offset, out.data(),
                                                     out.size(), S252_HOST));
        return out;
    }

  private:
    std::vector<FE> coeffs_;
};

inline std::vector<FE> evaluate_polynomial_on_lde_domain(const Context& ctx, const Polynomial& p, size_t blowup_factor,
                                                         size_t domain_size, const FE& offset) {
    std::vector<FE> out(domain_size * blowup_factor);
    ctx.check<FFTError>(s252_evaluate_polynomial_on_lde_domain(ctx.raw(), p.coefficients().data(), p.coeff_len(), blowup_factor,
                                                               domain_size, &offset, out.data(), S252_HOST));
    return out;
}

struct Proof { std::vector<Commitment> merkle_path; };   // lambdaworks_crypto::merkle_tree::proof::Proof

// Device-resident commit: trace polynomials + LDE columns + batched Merkle tree.
class Commit {
  public:
    Commit() = default;
    Commit(const Context* ctx, s252_commit* h, const Commitment& r) : ctx_(ctx), h_(h), root(r) {}
    Commit(Commit&& o) noexcept : ctx_(o.ctx_), h_(o.h_), root(o.root) { o.h_ = nullptr; }
    Commit& operator=(Commit&& o) noexcept { if (this != &o) { reset(); ctx_ = o.ctx_; h_ = o.h_; root = o.root; o.h_ = nullptr; } return *this; }
    ~Commit() { reset(); }
    size_t n_rows() const { return s252_commit_n_rows(h_); }
    size_t n_cols() const { return s252_commit_n_cols(h_); }
    std::optional<Proof> get_proof_by_pos(size_t pos) const {
        if (pos >= n_rows()) return std::nullopt;
        size_t depth = 0;
        while (((size_t)1 << depth) < n_rows()) ++depth;
        Proof p;
        p.merkle_path.resize(depth);
        uint64_t idx = pos;
        ctx_->check(s252_commit_open(h_, &idx, 1, nullptr, depth ? p.merkle_path[0].data() : nullptr));
        return p;
    }
    std::vector<FE> lde_column(size_t col) const {
        std::vector<FE> out(n_rows());
        ctx_->check(s252_commit_read_lde(h_, col, 0, out.size(), out.data()));
        return out;
    }
    Polynomial trace_poly(size_t col) const {
        std::vector<FE> out(s252_commit_n_coeffs(h_));
        ctx_->check(s252_commit_read_coeffs(h_, col, out.data()));
        return Polynomial(std::move(out));
    }
    s252_commit* raw() const { return h_; }
    Commitment root{};

  private:
    void reset() { if (h_) s252_commit_destroy(h_); h_ = nullptr; }
    const Context* ctx_ = nullptr;
    s252_commit* h_ = nullptr;
};

struct BatchedMerkleTree {
    // rows: row-major n_rows x n_cols
    static Commit build(const Context& ctx, const std::vector<FE>& rows, size_t n_cols) {
        s252_commit* h = nullptr;
        Commitment root;
        ctx.check(s252_merkle_build(ctx.raw(), rows.data(), n_cols ? rows.size() / n_cols : 0, n_cols, S252_HOST, &h, root.data()));
        return Commit(&ctx, h, root);
    }
};

struct TraceTable {                              // src/starks/trace.rs:9-13
    std::vector<FE> table;
    size_t n_cols = 0;
    size_t n_rows() const { return n_cols ? table.size() / n_cols : 0; }
};

// src/starks/prover.rs:126-159; appends the root to the transcript like the reference (prover.rs:151).
inline Commit interpolate_and_commit(const Context& ctx, const TraceTable& trace, const ProofOptions& options,
                                     DefaultTranscript& transcript) {
    s252_commit* h = nullptr;
    Commitment root;
    ctx.check<FFTError>(s252_interpolate_and_commit(ctx.raw(), trace.table.data(), trace.n_rows(), trace.n_cols,
                                                    options.blowup_factor, options.coset_offset, S252_HOST, &h, root.data()));
    transcript.append(root);
    return Commit(&ctx, h, root);
}

class FriLayers {
  public:
    FriLayers(const Context* ctx, s252_fri* h, std::vector<Commitment> roots) : ctx_(ctx), h_(h), roots(std::move(roots)) {}
    FriLayers(FriLayers&& o) noexcept : ctx_(o.ctx_), h_(o.h_), roots(std::move(o.roots)) { o.h_ = nullptr; }
    ~FriLayers() { if (h_) s252_fri_destroy(h_); }
    size_t len() const { return roots.size(); }
    std::vector<FE> evaluation(size_t layer, size_t domain_size) const {
        std::vector<FE> out(domain_size >> layer);
        ctx_->check(s252_fri_read_layer(h_, layer, 0, out.size(), out.data()));
        return out;
    }
    s252_fri* raw() const { return h_; }

  private:
    const Context* ctx_;
    s252_fri* h_;

  public:
    std::vector<Commitment> roots;               // layer.merkle_tree.root for every layer
};

// src/starks/fri/mod.rs:20-72 -> (last_value, fri_layer_list)
inline std::pair<FE, FriLayers> fri_commit_phase(const Context& ctx, size_t number_layers, const Polynomial& p_0,
                                                 DefaultTranscript& transcript, const FE& coset_offset, size_t domain_size) {
    s252_fri* h = nullptr;
    FE last;
    std::vector<Commitment> roots(number_layers);
    ctx.check(s252_fri_commit_phase(ctx.raw(), number_layers, p_0.coefficients().data(), p_0.coeff_len(), transcript.raw(),
                                    &coset_offset, domain_size, S252_HOST, &h, &last,
                                    number_layers ? roots[0].data() : nullptr));
    return {last, FriLayers(&ctx, h, std::move(roots))};
}

// src/starks/grinding.rs:40-48 -> Option<u64>
inline std::optional<uint64_t> generate_nonce_with_grinding(const Context& ctx, const Commitment& transcript_challenge,
                                                            uint8_t grinding_factor) {
    uint64_t nonce = 0;
    int rc = s252_generate_nonce_with_grinding(ctx.raw(), transcript_challenge.data(), grinding_factor, 0, &nonce);
    if (rc == S252_ERR_NOT_FOUND) return std::nullopt;
    ctx.check(rc);
    return nonce;
}

}  // namespace stark252
