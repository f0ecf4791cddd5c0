// host_ff.hpp — host-side BN254 Fq / Fr (4x64-bit limbs, Montgomery, fully reduced).
//
// Product host code (proof assembly glue, synthetic proving-key scalars, witness generation, the
// off-chain pairing verifier).  This is NOT a CPU fallback for the GPU hot path: nothing here
// computes an MSM or an NTT.  The reference has no counterpart (SURVEY.md finding 1); the
// consumer of what this code verifies is EigenZkVM.verifyBatches
// (reference src/settlement/ethereum/interfaces/zkvm.rs:82-130).
#pragma once
#include <cstdint>
#include <cstring>

namespace zkb {
namespace host {

typedef unsigned __int128 u128;

struct FqCfg {
  static constexpr uint64_t MOD[4] = {0x3c208c16d87cfd47ull, 0x97816a916871ca8dull,
                                      0xb85045b68181585dull, 0x30644e72e131a029ull};
  static constexpr uint64_t ONE[4] = {0xd35d438dc58f0d9dull, 0x0a78eb28f5c70b3dull,
                                      0x666ea36f7879462cull, 0x0e0a77c19a07df2full};
  static constexpr uint64_t R2[4] = {0xf32cfc5b538afa89ull, 0xb5e71911d44501fbull,
                                     0x47ab1eff0a417ff6ull, 0x06d89f71cab8351full};
  static constexpr uint64_t INV = 0x87d20782e4866389ull;
};
struct FrCfg {
  static constexpr uint64_t MOD[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull,
                                      0xb85045b68181585dull, 0x30644e72e131a029ull};
  static constexpr uint64_t ONE[4] = {0xac96341c4ffffffbull, 0x36fc76959f60cd29ull,
                                      0x666ea36f7879462eull, 0x0e0a77c19a07df2full};
  static constexpr uint64_t R2[4] = {0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull,
                                     0x8c49833d53bb8085ull, 0x0216d0b17f4e44a5ull};
  static constexpr uint64_t INV = 0xc2e1f593efffffffull;
};

template <class C>
struct HFp {
  uint64_t l[4];

  static HFp zero() { return HFp{{0, 0, 0, 0}}; }
  static HFp one() { return HFp{{C::ONE[0], C::ONE[1], C::ONE[2], C::ONE[3]}}; }
  static HFp from_u64(uint64_t x) {
    HFp a{{x, 0, 0, 0}};
    return a.to_mont();
  }
  // canonical little-endian bytes -> Montgomery (value must be < p; returns false otherwise)
  static bool from_bytes(const uint8_t* b, HFp* out) {
    HFp a;
    std::memcpy(a.l, b, 32);
    if (!lt_mod(a.l)) return false;
    *out = a.to_mont();
    return true;
  }
  void to_bytes(uint8_t* b) const {
    HFp a = from_mont();
    std::memcpy(b, a.l, 32);
  }
  static bool lt_mod(const uint64_t* a) {
    for (int i = 3; i >= 0; i--) {
      if (a[i] < C::MOD[i]) return true;
      if (a[i] > C::MOD[i]) return false;
    }
    return false;
  }
  bool is_zero() const { return (l[0] | l[1] | l[2] | l[3]) == 0; }
  bool operator==(const HFp& o) const {
    return l[0] == o.l[0] && l[1] == o.l[1] && l[2] == o.l[2] && l[3] == o.l[3];
  }
  bool operator!=(const HFp& o) const { return !(*this == o); }

  HFp operator+(const HFp& o) const {
    HFp r;
    u128 c = 0;
    for (int i = 0; i < 4; i++) {
      c += (u128)l[i] + o.l[i];
      r.l[i] = (uint64_t)c;
      c >>= 64;
    }
    r.cond_sub();
    return r;
  }
  HFp operator-(const HFp& o) const {
    HFp r;
    uint64_t borrow = 0;
    for (int i = 0; i < 4; i++) {
      u128 d = (u128)l[i] - o.l[i] - borrow;
      r.l[i] = (uint64_t)d;
      borrow = (uint64_t)(d >> 64) & 1;
    }
    if (borrow) {
      u128 c = 0;
      for (int i = 0; i < 4; i++) {
        c += (u128)r.l[i] + C::MOD[i];
        r.l[i] = (uint64_t)c;
        c >>= 64;
      }
    }
    return r;
  }
  HFp neg() const { return zero() - *this; }
  HFp dbl() const { return *this + *this; }

  HFp operator*(const HFp& o) const {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
      u128 c = 0;
      for (int j = 0; j < 4; j++) {
        c += (u128)l[j] * o.l[i] + t[j];
        t[j] = (uint64_t)c;
        c >>= 64;
      }
      c += t[4];
      t[4] = (uint64_t)c;
      t[5] = (uint64_t)(c >> 64);
      uint64_t m = t[0] * C::INV;
      c = (u128)m * C::MOD[0] + t[0];
      c >>= 64;
      for (int j = 1; j < 4; j++) {
        c += (u128)m * C::MOD[j] + t[j];
        t[j - 1] = (uint64_t)c;
        c >>= 64;
      }
      c += t[4];
      t[3] = (uint64_t)c;
      t[4] = t[5] + (uint64_t)(c >> 64);
    }
    HFp r{{t[0], t[1], t[2], t[3]}};
    if (t[4] || !lt_mod(r.l)) r.raw_sub_mod();
    return r;
  }
  HFp sqr() const { return *this * *this; }
  HFp to_mont() const { return *this * HFp{{C::R2[0], C::R2[1], C::R2[2], C::R2[3]}}; }
  HFp from_mont() const { return *this * HFp{{1, 0, 0, 0}}; }

  HFp pow(const uint64_t* e, int nwords) const {
    HFp r = one();
    for (int i = nwords * 64 - 1; i >= 0; i--) {
      r = r.sqr();
      if ((e[i >> 6] >> (i & 63)) & 1) r = r * *this;
    }
    return r;
  }
  HFp inverse() const {  // Fermat; 0 -> 0
    uint64_t e[4] = {C::MOD[0] - 2, C::MOD[1], C::MOD[2], C::MOD[3]};
    return pow(e, 4);
  }

 private:
  void raw_sub_mod() {
    uint64_t borrow = 0;
    for (int i = 0; i < 4; i++) {
      u128 d = (u128)l[i] - C::MOD[i] - borrow;
      l[i] = (uint64_t)d;
      borrow = (uint64_t)(d >> 64) & 1;
    }
  }
  void cond_sub() {
    if (!lt_mod(l)) raw_sub_mod();
  }
};

using HFq = HFp<FqCfg>;
using HFr = HFp<FrCfg>;

// xoshiro256** seeded through splitmix64 — the synthetic-workload generator shared with the oracles
struct Xoshiro256ss {
  uint64_t s[4];
  explicit Xoshiro256ss(uint64_t seed) {
    uint64_t x = seed;
    for (int i = 0; i < 4; i++) {
      x += 0x9E3779B97F4A7C15ull;
      uint64_t z = x;
      z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
      z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
      s[i] = z ^ (z >> 31);
    }
  }
  static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
  uint64_t next() {
    uint64_t result = rotl(s[1] * 5, 7) * 9;
    uint64_t t = s[1] << 17;
    s[2] ^= s[0];
    s[3] ^= s[1];
    s[1] ^= s[2];
    s[0] ^= s[3];
    s[2] ^= t;
    s[3] = rotl(s[3], 45);
    return result;
  }
  // uniform canonical Fr in [0, r): 254-bit mask + rejection, little-endian limbs
  void fr_canonical(uint64_t out[4]) {
    for (;;) {
      for (int i = 0; i < 4; i++) out[i] = next();
      out[3] &= (1ull << 62) - 1;
      if (HFr::lt_mod(out)) return;
    }
  }
  HFr fr() {
    uint64_t v[4];
    fr_canonical(v);
    HFr a{{v[0], v[1], v[2], v[3]}};
    return a.to_mont();
  }
};

}  // namespace host
}  // namespace zkb
