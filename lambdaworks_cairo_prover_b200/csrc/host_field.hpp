// host_field.hpp -- host-side Stark252 scalars, Keccak-256 and the Fiat-Shamir transcript.
//
// The GPU does the bulk arithmetic; the host only needs a handful of scalars per call (roots of
// unity, coset shifts, zeta/(2h), 1/N) and the transcript, which is inherently sequential
// (every FRI challenge depends on the previous layer's root, src/starks/fri/mod.rs:37-54).
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

#include "fe.cuh"

namespace s252 {
namespace host {

typedef unsigned __int128 u128;

struct U256 { uint64_t w[4]; };   // little-endian words

inline U256 to_u256(const fe& a) {
    U256 r;
    for (int i = 0; i < 4; ++i) r.w[i] = ((uint64_t)a.l[2 * i + 1] << 32) | a.l[2 * i];
    return r;
}
inline fe from_u256(const U256& a) {
    fe r;
    for (int i = 0; i < 4; ++i) { r.l[2 * i] = (uint32_t)a.w[i]; r.l[2 * i + 1] = (uint32_t)(a.w[i] >> 32); }
    return r;
}
static const U256 MOD = {{1ULL, 0ULL, 0ULL, 0x0800000000000011ULL}};

inline bool geq_mod(const U256& a) {
    for (int i = 3; i >= 0; --i) {
        if (a.w[i] != MOD.w[i]) return a.w[i] > MOD.w[i];
    }
    return true;
}
inline void sub_mod_inplace(U256& a) {
    u128 borrow = 0;
    for (int i = 0; i < 4; ++i) {
        u128 d = (u128)a.w[i] - MOD.w[i] - borrow;
        a.w[i] = (uint64_t)d;
        borrow = (d >> 127) & 1;
    }
}
// Montgomery product, word-serial with the quotient digit -t0 (p = 1 mod 2^64).
inline fe mul(const fe& x, const fe& y) {
    const U256 a = to_u256(x), b = to_u256(y);
    uint64_t t[5] = {0, 0, 0, 0, 0};
    uint64_t t5 = 0;
    for (int i = 0; i < 4; ++i) {
        u128 carry = 0;
        for (int j = 0; j < 4; ++j) {
            u128 cur = (u128)a.w[j] * b.w[i] + t[j] + carry;
            t[j] = (uint64_t)cur;
            carry = cur >> 64;
        }
        u128 top = (u128)t[4] + carry;
        t[4] = (uint64_t)top;
        t5 = (uint64_t)(top >> 64);
        const uint64_t m = 0 - t[0];
        // t += m * p ; p = 1 + MOD.w[3] * 2^192 ; then shift one word down
        u128 cur = (u128)t[0] + m;                 // low word becomes zero
        carry = cur >> 64;
        for (int j = 1; j < 3; ++j) { cur = (u128)t[j] + carry; t[j - 1] = (uint64_t)cur; carry = cur >> 64; }
        cur = (u128)m * MOD.w[3] + t[3] + carry;
        t[2] = (uint64_t)cur;
        carry = cur >> 64;
        cur = (u128)t[4] + carry;
        t[3] = (uint64_t)cur;
        t[4] = t5 + (uint64_t)(cur >> 64);
    }
    U256 r = {{t[0], t[1], t[2], t[3]}};
    if (t[4] || geq_mod(r)) sub_mod_inplace(r);
    return from_u256(r);
}
inline fe add(const fe& x, const fe& y) {
    const U256 a = to_u256(x), b = to_u256(y);
    U256 r;
    u128 carry = 0;
    for (int i = 0; i < 4; ++i) { u128 s = (u128)a.w[i] + b.w[i] + carry; r.w[i] = (uint64_t)s; carry = s >> 64; }
    if (geq_mod(r)) sub_mod_inplace(r);
    return from_u256(r);
}
inline fe sub(const fe& x, const fe& y) {
    const U256 a = to_u256(x), b = to_u256(y);
    U256 r;
    u128 borrow = 0;
    for (int i = 0; i < 4; ++i) {
        u128 d = (u128)a.w[i] - b.w[i] - borrow;
        r.w[i] = (uint64_t)d;
        borrow = (d >> 127) & 1;
    }
    if (borrow) {
        u128 carry = 0;
        for (int i = 0; i < 4; ++i) { u128 s = (u128)r.w[i] + MOD.w[i] + carry; r.w[i] = (uint64_t)s; carry = s >> 64; }
    }
    return from_u256(r);
}
inline fe one() { return fe_one(); }
inline fe zero() { return fe_zero(); }
inline bool is_zero(const fe& a) {
    uint32_t o = 0;
    for (int i = 0; i < 8; ++i) o |= a.l[i];
    return o == 0;
}
inline bool eq(const fe& a, const fe& b) { return std::memcmp(a.l, b.l, 32) == 0; }
inline fe to_mont(const fe& canon) { return mul(canon, fe_r2()); }
inline fe from_mont(const fe& m) { fe o = fe{{1, 0, 0, 0, 0, 0, 0, 0}}; return mul(m, o); }
inline fe from_u64(uint64_t v) { fe c = fe{{(uint32_t)v, (uint32_t)(v >> 32), 0, 0, 0, 0, 0, 0}}; return to_mont(c); }
inline fe sqr(const fe& a) { return mul(a, a); }
inline fe pow_u64(fe a, uint64_t e) {
    fe r = one();
    while (e) { if (e & 1) r = mul(r, a); a = sqr(a); e >>= 1; }
    return r;
}
inline fe inv(const fe& a) {
    // a^(p-2);  p - 2 = 0x0800000000000010 ffffffffffffffff ffffffffffffffff ffffffffffffffff
    const uint64_t e[4] = {0xffffffffffffffffULL, 0xffffffffffffffffULL, 0xffffffffffffffffULL, 0x0800000000000010ULL};
    fe r = one(), b = a;
    for (int i = 0; i < 4; ++i)
        for (int k = 0; k < 64; ++k) {
            if ((e[i] >> k) & 1) r = mul(r, b);
            b = sqr(b);
        }
    return r;
}
// F::get_primitive_root_of_unity(order) = W^(2^(192-order))  (call sites domain.rs:30, verifier.rs:367)
inline bool primitive_root(unsigned order, fe* out) {
    if (order > 192) return false;
    const fe w_canon = fe{{0x42f8ef94u, 0x6070024fu, 0xe11a6161u, 0xad187148u, 0x9c8b0fa5u, 0x3f046451u, 0x87529cfau, 0x005282dbu}};
    fe w = to_mont(w_canon);
    for (unsigned i = 0; i < 192 - order; ++i) w = sqr(w);
    *out = w;
    return true;
}
// LW (4 x u64, most significant first) <-> internal
inline fe from_lw(const uint64_t limbs[4]) {
    U256 v = {{limbs[3], limbs[2], limbs[1], limbs[0]}};
    return from_u256(v);
}
inline void to_lw(const fe& a, uint64_t limbs[4]) {
    const U256 v = to_u256(a);
    limbs[0] = v.w[3]; limbs[1] = v.w[2]; limbs[2] = v.w[1]; limbs[3] = v.w[0];
}
// ByteConversion::to_bytes_be / from_bytes_be (canonical value, 32 bytes big-endian)
inline void to_bytes_be(const fe& m, uint8_t out[32]) {
    const U256 c = to_u256(from_mont(m));
    for (int i = 0; i < 4; ++i)
        for (int k = 0; k < 8; ++k) out[8 * i + k] = (uint8_t)(c.w[3 - i] >> (56 - 8 * k));
}
inline fe from_bytes_be(const uint8_t in[32]) {
    U256 c;
    for (int i = 0; i < 4; ++i) {
        uint64_t v = 0;
        for (int k = 0; k < 8; ++k) v = (v << 8) | in[8 * i + k];
        c.w[3 - i] = v;
    }
    return to_mont(from_u256(c));
}

// ---- Keccak-256 sponge (sha3 crate's Keccak256: pad 0x01, rate 136) -----------------------
class Keccak256 {
  public:
    Keccak256() { reset(); }
    void reset() { std::memset(s_, 0, sizeof s_); fill_ = 0; }
    void update(const uint8_t* d, size_t n) {
        while (n) {
            size_t take = 136 - fill_;
            if (take > n) take = n;
            for (size_t i = 0; i < take; ++i) {
                const size_t pos = fill_ + i;
                s_[pos >> 3] ^= (uint64_t)d[i] << (8 * (pos & 7));
            }
            fill_ += take; d += take; n -= take;
            if (fill_ == 136) { permute(); fill_ = 0; }
        }
    }
    // the sponge as it stands (absorbed bytes are XORed into the lanes): what a device-side continuation needs
    const uint64_t* lanes() const { return s_; }
    size_t fill() const { return fill_; }
    void finalize(uint8_t out[32]) {
        s_[fill_ >> 3] ^= (uint64_t)0x01 << (8 * (fill_ & 7));
        s_[16] ^= 0x8000000000000000ULL;
        permute();
        for (int i = 0; i < 32; ++i) out[i] = (uint8_t)(s_[i >> 3] >> (8 * (i & 7)));
    }

  private:
    static uint64_t rotl(uint64_t x, int n) { return n ? (x << n) | (x >> (64 - n)) : x; }
    void permute() {
        static const uint64_t RC[24] = {
            0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
            0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
            0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
            0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
            0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
            0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
        // rho offsets walked along the pi cycle starting from lane 1
        static const int PI[24] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4, 15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};
        static const int RHO[24] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14, 27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
        for (int round = 0; round < 24; ++round) {
            uint64_t c[5];
            for (int x = 0; x < 5; ++x) c[x] = s_[x] ^ s_[x + 5] ^ s_[x + 10] ^ s_[x + 15] ^ s_[x + 20];
            for (int x = 0; x < 5; ++x) {
                const uint64_t d = c[(x + 4) % 5] ^ rotl(c[(x + 1) % 5], 1);
                for (int y = 0; y < 25; y += 5) s_[y + x] ^= d;
            }
            uint64_t cur = s_[1];
            for (int i = 0; i < 24; ++i) {
                const int j = PI[i];
                const uint64_t tmp = s_[j];
                s_[j] = rotl(cur, RHO[i]);
                cur = tmp;
            }
            for (int y = 0; y < 25; y += 5) {
                uint64_t row[5];
                for (int x = 0; x < 5; ++x) row[x] = s_[y + x];
                for (int x = 0; x < 5; ++x) s_[y + x] = row[x] ^ (~row[(x + 1) % 5] & row[(x + 2) % 5]);
            }
            s_[0] ^= RC[round];
        }
    }
    uint64_t s_[25];
    size_t fill_;
};

}  // namespace host
}  // namespace s252

// DefaultTranscript: append = absorb; challenge = finalize, reverse the 32 bytes, reset, absorb them.
struct s252_transcript {
    s252::host::Keccak256 k;
    void append(const uint8_t* d, size_t n) { k.update(d, n); }
    void challenge(uint8_t out[32]) {
        uint8_t h[32];
        k.finalize(h);
        for (int i = 0; i < 32; ++i) out[i] = h[31 - i];
        k.reset();
        k.update(out, 32);
    }
    // src/starks/transcript.rs:13-43 (251 random bits, big-endian)
    s252::fe to_field() {
        uint8_t r[32];
        challenge(r);
        r[0] &= 0x07;
        return s252::host::from_bytes_be(r);
    }
    // src/starks/transcript.rs:45-51
    uint64_t to_usize() {
        uint8_t r[32];
        challenge(r);
        uint64_t v = 0;
        for (int i = 0; i < 8; ++i) v = (v << 8) | r[i];
        return v;
    }
};
