// oracle_b.cpp — Oracle B: C++17/OpenMP CPU restatement of the BN254 Groth16 proving path.
//
// TEST INFRASTRUCTURE AND TIMED CPU BASELINE ONLY.  Nothing in the product (eigen-zeth_b200/)
// may include, link or call this file; only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs do.
//
// What it restates.  The reference (eigen-zeth) holds no proving arithmetic: the work is done by
// an external service behind GenFinalProofRequest (reference proto/prover/v1/prover.proto:130-148,
// call site src/prover/provider.rs:472-483), and the north_star names arkworks as the CPU
// baseline.  arkworks is absent from the mount and cannot be built here (no Rust toolchain), so
// this file restates its published algorithms (arkworks 0.4):
//   * ark-ff Montgomery fields on 4x64-bit limbs,
//   * ark-ec short-Weierstrass Jacobian arithmetic and VariableBaseMSM::msm_bigint
//     (window c = 3 if n < 32 else ceil(ln n) + 2, unsigned digits, one bucket set per window,
//     windows in parallel, running-sum bucket reduction, Horner over windows),
//   * ark-poly Radix2EvaluationDomain fft/ifft/coset (in-place radix-2, bit reversal, g = 5),
//   * ark-groth16 create_proof_with_reduction_and_matrices + LibsnarkReduction.
// PARITY UNPINNED against the reference itself (it has no known-answer vectors for proof values,
// SURVEY.md Appendix C); this oracle is pinned bit-for-bit against Oracle A (oracle_a.py, pure
// Python integers) by tests/test_oracles.py, and by the Groth16 pairing equation.
//
// All byte interfaces: 32-byte little-endian canonical field elements; G1 affine x||y (64 B);
// G2 affine x.c0||x.c1||y.c0||y.c1 (128 B); all-zero = point at infinity.
#include <omp.h>

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

typedef unsigned __int128 u128;

namespace {

// ------------------------------------------------------------------ 256-bit modular arithmetic
struct U256 {
  uint64_t w[4];
};
struct Modulus {
  U256 p;       // modulus
  U256 r1;      // 2^256 mod p   (Montgomery one)
  U256 r2;      // 2^512 mod p
  uint64_t ninv;  // -p^{-1} mod 2^64
};

const Modulus FQ = {{{0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull}},
                    {{0xd35d438dc58f0d9dull, 0x0a78eb28f5c70b3dull, 0x666ea36f7879462cull, 0x0e0a77c19a07df2full}},
                    {{0xf32cfc5b538afa89ull, 0xb5e71911d44501fbull, 0x47ab1eff0a417ff6ull, 0x06d89f71cab8351full}},
                    0x87d20782e4866389ull};
const Modulus FR = {{{0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull}},
                    {{0xac96341c4ffffffbull, 0x36fc76959f60cd29ull, 0x666ea36f7879462eull, 0x0e0a77c19a07df2full}},
                    {{0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull, 0x8c49833d53bb8085ull, 0x0216d0b17f4e44a5ull}},
                    0xc2e1f593efffffffull};

inline bool geq(const U256& a, const U256& b) {
  for (int i = 3; i >= 0; i--) {
    if (a.w[i] != b.w[i]) return a.w[i] > b.w[i];
  }
  return true;
}
inline uint64_t add_to(U256& a, const U256& b) {
  u128 c = 0;
  for (int i = 0; i < 4; i++) {
    c += (u128)a.w[i] + b.w[i];
    a.w[i] = (uint64_t)c;
    c >>= 64;
  }
  return (uint64_t)c;
}
inline uint64_t sub_from(U256& a, const U256& b) {
  uint64_t br = 0;
  for (int i = 0; i < 4; i++) {
    u128 d = (u128)a.w[i] - b.w[i] - br;
    a.w[i] = (uint64_t)d;
    br = (uint64_t)(d >> 127);
  }
  return br;
}

// separated-operand-scanning Montgomery product (different schedule from the product's CIOS)
inline U256 mmul(const U256& a, const U256& b, const Modulus& M) {
  uint64_t t[9] = {0};
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) {
      c += (u128)a.w[i] * b.w[j] + t[i + j];
      t[i + j] = (uint64_t)c;
      c >>= 64;
    }
    t[i + 4] = (uint64_t)c;
  }
  for (int i = 0; i < 4; i++) {
    uint64_t m = t[i] * M.ninv;
    u128 c = 0;
    for (int j = 0; j < 4; j++) {
      c += (u128)m * M.p.w[j] + t[i + j];
      t[i + j] = (uint64_t)c;
      c >>= 64;
    }
    for (int k = i + 4; c && k < 9; k++) {
      c += t[k];
      t[k] = (uint64_t)c;
      c >>= 64;
    }
  }
  U256 r = {{t[4], t[5], t[6], t[7]}};
  if (t[8] || geq(r, M.p)) sub_from(r, M.p);
  return r;
}
inline U256 madd(const U256& a, const U256& b, const Modulus& M) {
  U256 r = a;
  uint64_t c = add_to(r, b);
  if (c || geq(r, M.p)) sub_from(r, M.p);
  return r;
}
inline U256 msub(const U256& a, const U256& b, const Modulus& M) {
  U256 r = a;
  if (sub_from(r, b)) add_to(r, M.p);
  return r;
}
inline bool is0(const U256& a) { return (a.w[0] | a.w[1] | a.w[2] | a.w[3]) == 0; }
inline bool eq(const U256& a, const U256& b) {
  return a.w[0] == b.w[0] && a.w[1] == b.w[1] && a.w[2] == b.w[2] && a.w[3] == b.w[3];
}
const U256 ZERO = {{0, 0, 0, 0}};
inline U256 to_m(const U256& a, const Modulus& M) { return mmul(a, M.r2, M); }
inline U256 from_m(const U256& a, const Modulus& M) {
  U256 one = {{1, 0, 0, 0}};
  return mmul(a, one, M);
}
U256 mpow(U256 b, const uint64_t* e, int words, const Modulus& M) {
  U256 r = M.r1;
  for (int i = words * 64 - 1; i >= 0; i--) {
    r = mmul(r, r, M);
    if ((e[i >> 6] >> (i & 63)) & 1) r = mmul(r, b, M);
  }
  return r;
}
U256 minv(const U256& a, const Modulus& M) {
  U256 e = M.p;
  e.w[0] -= 2;
  return mpow(a, e.w, 4, M);
}
U256 mpow64(const U256& b, uint64_t e, const Modulus& M) { return mpow(b, &e, 1, M); }
U256 from_u64(uint64_t x, const Modulus& M) {
  U256 a = {{x, 0, 0, 0}};
  return to_m(a, M);
}

// ------------------------------------------------------------------ field wrappers for generic curves
struct Fq {
  U256 v;
  static Fq zero() { return {ZERO}; }
  static Fq one() { return {FQ.r1}; }
  Fq operator+(const Fq& o) const { return {madd(v, o.v, FQ)}; }
  Fq operator-(const Fq& o) const { return {msub(v, o.v, FQ)}; }
  Fq operator*(const Fq& o) const { return {mmul(v, o.v, FQ)}; }
  Fq sq() const { return {mmul(v, v, FQ)}; }
  Fq dbl() const { return {madd(v, v, FQ)}; }
  Fq neg() const { return {msub(ZERO, v, FQ)}; }
  Fq inv() const { return {minv(v, FQ)}; }
  bool zero_p() const { return is0(v); }
  bool operator==(const Fq& o) const { return eq(v, o.v); }
  static const int BYTES = 32;
  void put(uint8_t* b) const {
    U256 c = from_m(v, FQ);
    memcpy(b, c.w, 32);
  }
  static Fq get(const uint8_t* b) {
    U256 c;
    memcpy(c.w, b, 32);
    return {to_m(c, FQ)};
  }
};
struct Fq2 {
  Fq a, b;  // a + b u, u^2 = -1
  static Fq2 zero() { return {Fq::zero(), Fq::zero()}; }
  static Fq2 one() { return {Fq::one(), Fq::zero()}; }
  Fq2 operator+(const Fq2& o) const { return {a + o.a, b + o.b}; }
  Fq2 operator-(const Fq2& o) const { return {a - o.a, b - o.b}; }
  Fq2 operator*(const Fq2& o) const {  // schoolbook (4 mults): deliberately not the product's Karatsuba
    return {a * o.a - b * o.b, a * o.b + b * o.a};
  }
  Fq2 sq() const { return *this * *this; }
  Fq2 dbl() const { return {a.dbl(), b.dbl()}; }
  Fq2 neg() const { return {a.neg(), b.neg()}; }
  Fq2 inv() const {
    Fq d = (a.sq() + b.sq()).inv();
    return {a * d, (b * d).neg()};
  }
  bool zero_p() const { return a.zero_p() && b.zero_p(); }
  bool operator==(const Fq2& o) const { return a == o.a && b == o.b; }
  static const int BYTES = 64;
  void put(uint8_t* p) const {
    a.put(p);
    b.put(p + 32);
  }
  static Fq2 get(const uint8_t* p) { return {Fq::get(p), Fq::get(p + 32)}; }
};

// ------------------------------------------------------------------ Jacobian short Weierstrass, a = 0
template <class F>
struct Aff {
  F x, y;
  bool inf;
};
template <class F>
struct Jac {
  F x, y, z;
  static Jac identity() { return {F::one(), F::one(), F::zero()}; }
  bool is_identity() const { return z.zero_p(); }
  // dbl-2009-l
  Jac dbl() const {
    if (is_identity()) return *this;
    F A = x.sq(), B = y.sq(), C = B.sq();
    F D = ((x + B).sq() - A - C).dbl();
    F E = A.dbl() + A, Fv = E.sq();
    Jac r;
    r.x = Fv - D.dbl();
    r.z = (y * z).dbl();
    r.y = E * (D - r.x) - C.dbl().dbl().dbl();
    return r;
  }
  // madd-2007-bl
  void add_affine(const Aff<F>& q) {
    if (q.inf) return;
    if (is_identity()) {
      x = q.x; y = q.y; z = F::one();
      return;
    }
    F Z1Z1 = z.sq();
    F U2 = q.x * Z1Z1;
    F S2 = q.y * z * Z1Z1;
    if (U2 == x) {
      if (S2 == y) *this = dbl();
      else *this = identity();
      return;
    }
    F H = U2 - x, HH = H.sq(), I = HH.dbl().dbl(), J = H * I;
    F rr = (S2 - y).dbl();
    F V = x * I;
    Jac r;
    r.x = rr.sq() - J - V.dbl();
    r.y = rr * (V - r.x) - (y * J).dbl();
    r.z = (z + H).sq() - Z1Z1 - HH;
    *this = r;
  }
  // add-2007-bl
  void add(const Jac& q) {
    if (q.is_identity()) return;
    if (is_identity()) {
      *this = q;
      return;
    }
    F Z1Z1 = z.sq(), Z2Z2 = q.z.sq();
    F U1 = x * Z2Z2, U2 = q.x * Z1Z1;
    F S1 = y * q.z * Z2Z2, S2 = q.y * z * Z1Z1;
    if (U1 == U2) {
      if (S1 == S2) *this = dbl();
      else *this = identity();
      return;
    }
    F H = U2 - U1, I = H.dbl().sq(), J = H * I;
    F rr = (S2 - S1).dbl();
    F V = U1 * I;
    Jac r;
    r.x = rr.sq() - J - V.dbl();
    r.y = rr * (V - r.x) - (S1 * J).dbl();
    r.z = ((z + q.z).sq() - Z1Z1 - Z2Z2) * H;
    *this = r;
  }
  Aff<F> affine() const {
    if (is_identity()) return {F::zero(), F::zero(), true};
    F zi = z.inv(), zi2 = zi.sq();
    return {x * zi2, y * zi2 * zi, false};
  }
};

template <class F>
Aff<F> aff_get(const uint8_t* b) {
  bool z = true;
  for (int i = 0; i < 2 * F::BYTES; i++) z = z && b[i] == 0;
  if (z) return {F::zero(), F::zero(), true};
  return {F::get(b), F::get(b + F::BYTES), false};
}
template <class F>
void aff_put(const Aff<F>& p, uint8_t* b) {
  if (p.inf) {
    memset(b, 0, 2 * F::BYTES);
    return;
  }
  p.x.put(b);
  p.y.put(b + F::BYTES);
}
template <class F>
Jac<F> scalar_mul(const Aff<F>& p, const U256& k) {
  Jac<F> r = Jac<F>::identity();
  for (int i = 255; i >= 0; i--) {
    r = r.dbl();
    if ((k.w[i >> 6] >> (i & 63)) & 1) r.add_affine(p);
  }
  return r;
}

const Aff<Fq> G1GEN = {{from_u64(1, FQ)}, {from_u64(2, FQ)}, false};
Aff<Fq2> g2gen() {
  static const uint64_t c[4][4] = {
      {0x46debd5cd992f6edull, 0x674322d4f75edaddull, 0x426a00665e5c4479ull, 0x1800deef121f1e76ull},
      {0x97e485b7aef312c2ull, 0xf1aa493335a9e712ull, 0x7260bfb731fb5d25ull, 0x198e9393920d483aull},
      {0x4ce6cc0166fa7daaull, 0xe3d1e7690c43d37bull, 0x4aab71808dcb408full, 0x12c85ea5db8c6debull},
      {0x55acdadcd122975bull, 0xbc4b313370b38ef3ull, 0xec9e99ad690c3395ull, 0x090689d0585ff075ull}};
  Fq v[4];
  for (int i = 0; i < 4; i++) {
    U256 t = {{c[i][0], c[i][1], c[i][2], c[i][3]}};
    v[i] = {to_m(t, FQ)};
  }
  return {{v[0], v[1]}, {v[2], v[3]}, false};
}

// ------------------------------------------------------------------ VariableBaseMSM (ark-ec msm_bigint)
template <class F>
Jac<F> msm_arkworks(const Aff<F>* bases, const U256* scalars, size_t n) {
  if (n == 0) return Jac<F>::identity();
  size_t c = n < 32 ? 3 : (size_t)std::ceil(std::log((double)n)) + 2;
  const size_t num_bits = 254;
  size_t nwin = (num_bits + c - 1) / c;
  std::vector<Jac<F>> window_sums(nwin);
#pragma omp parallel for schedule(dynamic, 1)
  for (long wi = 0; wi < (long)nwin; wi++) {
    size_t w_start = wi * c;
    Jac<F> res = Jac<F>::identity();
    std::vector<Jac<F>> buckets(((size_t)1 << c) - 1, Jac<F>::identity());
    for (size_t i = 0; i < n; i++) {
      const U256& s = scalars[i];
      if (is0(s)) continue;
      // ark-ec adds scalar == 1 straight into the first window's result
      if (s.w[0] == 1 && !s.w[1] && !s.w[2] && !s.w[3]) {
        if (w_start == 0) res.add_affine(bases[i]);
        continue;
      }
      size_t word = w_start >> 6, sh = w_start & 63;
      uint64_t d = s.w[word] >> sh;
      if (sh && word + 1 < 4) d |= s.w[word + 1] << (64 - sh);
      d &= ((uint64_t)1 << c) - 1;
      if (d) buckets[d - 1].add_affine(bases[i]);
    }
    Jac<F> running = Jac<F>::identity();
    for (size_t b = buckets.size(); b-- > 0;) {
      running.add(buckets[b]);
      res.add(running);
    }
    window_sums[wi] = res;
  }
  Jac<F> total = window_sums[nwin - 1];
  for (size_t wi = nwin - 1; wi-- > 0;) {
    for (size_t k = 0; k < c; k++) total = total.dbl();
    total.add(window_sums[wi]);
  }
  return total;
}

// fixed-base windowed multiplication table for k*G (setup / synthetic bases only)
template <class F>
struct FixedBase {
  static const int W = 8, NW = 32;
  std::vector<Aff<F>> table;  // NW * 255 entries
  explicit FixedBase(const Aff<F>& g) : table((size_t)NW * 255) {
    Jac<F> base = {g.x, g.y, F::one()};
    for (int w = 0; w < NW; w++) {
      Jac<F> acc = base;
      for (int d = 1; d <= 255; d++) {
        table[(size_t)w * 255 + d - 1] = acc.affine();
        acc.add(base);
      }
      base = acc;   // 256 * previous base
    }
  }
  Jac<F> mul(const U256& k) const {
    Jac<F> r = Jac<F>::identity();
    for (int w = 0; w < NW; w++) {
      unsigned d = (unsigned)(k.w[w >> 3] >> ((w & 7) * 8)) & 255u;
      if (d) r.add_affine(table[(size_t)w * 255 + d - 1]);
    }
    return r;
  }
};

// batch Jacobian -> affine bytes
template <class F>
void batch_put(const std::vector<Jac<F>>& pts, uint8_t* out) {
  const size_t n = pts.size();
#pragma omp parallel for schedule(static)
  for (long i = 0; i < (long)n; i++) aff_put(pts[i].affine(), out + (size_t)i * 2 * F::BYTES);
}

// ------------------------------------------------------------------ radix-2 NTT over Fr (ark-poly)
const U256 ROOT_2_28 = {{0x9bd61b6e725b19f0ull, 0x402d111e41112ed4ull, 0x00e0a7eb8ef62abcull, 0x2a3c09f0a58a7e85ull}};

U256 fr_root(unsigned log_n, bool inverse) {
  U256 w = to_m(ROOT_2_28, FR);
  for (unsigned i = log_n; i < 28; i++) w = mmul(w, w, FR);
  return inverse ? minv(w, FR) : w;
}

// in-place, natural order in and out, values in Montgomery form
void fr_ntt_inplace(std::vector<U256>& a, unsigned log_n, bool inverse) {
  const size_t n = (size_t)1 << log_n;
  for (size_t i = 0; i < n; i++) {   // bit reversal permutation
    size_t j = 0;
    for (unsigned b = 0; b < log_n; b++) j |= ((i >> b) & 1) << (log_n - 1 - b);
    if (i < j) std::swap(a[i], a[j]);
  }
  U256 wn = fr_root(log_n, inverse);
  // powers table w^0..w^{n/2-1} (ark-poly precomputes roots the same way)
  std::vector<U256> roots(n / 2 ? n / 2 : 1);
  roots[0] = FR.r1;
  for (size_t i = 1; i < n / 2; i++) roots[i] = mmul(roots[i - 1], wn, FR);
  for (size_t m = 1; m < n; m <<= 1) {
    size_t stride = n / (2 * m);
#pragma omp parallel for schedule(static) if (n >= 4096)
    for (long kk = 0; kk < (long)(n / (2 * m)); kk++) {
      size_t k = (size_t)kk * 2 * m;
      for (size_t j = 0; j < m; j++) {
        U256 t = mmul(a[k + j + m], roots[j * stride], FR);
        U256 u = a[k + j];
        a[k + j] = madd(u, t, FR);
        a[k + j + m] = msub(u, t, FR);
      }
    }
  }
  if (inverse) {
    U256 ninv = minv(from_u64(n, FR), FR);
#pragma omp parallel for schedule(static) if (n >= 4096)
    for (long i = 0; i < (long)n; i++) a[i] = mmul(a[i], ninv, FR);
  }
}
void fr_distribute_powers(std::vector<U256>& a, const U256& g) {
  // sequential per chunk so the result does not depend on the thread count
  const size_t n = a.size();
  int nt = omp_get_max_threads();
  size_t chunk = (n + nt - 1) / nt;
#pragma omp parallel for schedule(static, 1)
  for (int t = 0; t < nt; t++) {
    size_t lo = (size_t)t * chunk, hi = lo + chunk < n ? lo + chunk : n;
    if (lo >= hi) continue;
    U256 p = mpow64(g, lo, FR);
    for (size_t i = lo; i < hi; i++) {
      a[i] = mmul(a[i], p, FR);
      p = mmul(p, g, FR);
    }
  }
}
void fr_ntt_full(std::vector<U256>& a, unsigned log_n, bool inverse, bool coset) {
  U256 g = from_u64(5, FR);
  if (coset && !inverse) fr_distribute_powers(a, g);
  fr_ntt_inplace(a, log_n, inverse);
  if (coset && inverse) fr_distribute_powers(a, minv(g, FR));
}

// ------------------------------------------------------------------ PRNG shared with oracle A and the product
struct Xoshiro {
  uint64_t s[4];
  explicit Xoshiro(uint64_t seed) {
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
    uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return r;
  }
  U256 fr() {  // canonical
    for (;;) {
      U256 v;
      for (int i = 0; i < 4; i++) v.w[i] = next();
      v.w[3] &= ((uint64_t)1 << 62) - 1;
      if (!geq(v, FR.p)) return v;
    }
  }
};

// ------------------------------------------------------------------ synthetic chain R1CS
struct Entry {
  uint32_t col;
  uint32_t coeff;
};
struct R1CS {
  unsigned log_n;
  size_t m, ncons;
  std::vector<std::vector<Entry>> A, B, C;
};
R1CS synthetic_r1cs(unsigned log_n) {
  R1CS cs;
  cs.log_n = log_n;
  cs.m = (size_t)1 << log_n;
  size_t m = cs.m;
  for (size_t j = 0; j + 4 < m; j++) {
    cs.A.push_back({{0, (uint32_t)(j % 7 + 1)}, {(uint32_t)(j + 2), 1}});
    cs.B.push_back({{(uint32_t)(j + 3), 1}});
    cs.C.push_back({{(uint32_t)(j + 4), 1}});
  }
  cs.A.push_back({{(uint32_t)(m - 1), 1}});
  cs.B.push_back({{(uint32_t)(m - 2), 1}});
  cs.C.push_back({{1, 1}});
  cs.A.push_back({{1, 1}, {2, 1}});
  cs.B.push_back({{0, 1}});
  cs.C.push_back({{1, 1}, {2, 1}});
  cs.ncons = cs.A.size();
  return cs;
}
U256 row_dot(const std::vector<Entry>& row, const std::vector<U256>& z) {
  U256 acc = ZERO;
  for (const Entry& e : row) acc = madd(acc, mmul(from_u64(e.coeff, FR), z[e.col], FR), FR);
  return acc;
}

struct ProvingKey {
  unsigned log_n;
  size_t n, m;
  Aff<Fq> alpha_g1, beta_g1, delta_g1;
  Aff<Fq2> beta_g2, delta_g2, gamma_g2;
  std::vector<Aff<Fq>> a_query, b_g1_query, l_query, h_query, gamma_abc;
  std::vector<Aff<Fq2>> b_g2_query;
  R1CS cs;
};

template <class F>
std::vector<Aff<F>> fixed_mul_many(const FixedBase<F>& fb, const std::vector<U256>& ks_mont) {
  std::vector<Aff<F>> out(ks_mont.size());
#pragma omp parallel for schedule(dynamic, 256)
  for (long i = 0; i < (long)ks_mont.size(); i++) out[i] = fb.mul(from_m(ks_mont[i], FR)).affine();
  return out;
}

ProvingKey* setup_synthetic(unsigned log_n, uint64_t seed) {
  ProvingKey* pk = new ProvingKey();
  pk->log_n = log_n;
  pk->cs = synthetic_r1cs(log_n);
  const R1CS& cs = pk->cs;
  size_t n = (size_t)1 << log_n, m = cs.m;
  pk->n = n;
  pk->m = m;
  Xoshiro rng(seed);
  U256 tau = to_m(rng.fr(), FR), alpha = to_m(rng.fr(), FR), beta = to_m(rng.fr(), FR),
       gamma = to_m(rng.fr(), FR), delta = to_m(rng.fr(), FR);
  // Lagrange basis at tau: L_j = (tau^n - 1)/n * w^j / (tau - w^j)
  U256 w = fr_root(log_n, false);
  U256 zt = msub(mpow64(tau, n, FR), FR.r1, FR);
  U256 ninv = minv(from_u64(n, FR), FR);
  std::vector<U256> L(n), den(n);
  U256 wj = FR.r1;
  for (size_t j = 0; j < n; j++) {
    den[j] = msub(tau, wj, FR);
    L[j] = mmul(mmul(zt, ninv, FR), wj, FR);
    wj = mmul(wj, w, FR);
  }
  {  // batch inversion of den
    std::vector<U256> pre(n);
    U256 acc = FR.r1;
    for (size_t j = 0; j < n; j++) {
      pre[j] = acc;
      acc = mmul(acc, den[j], FR);
    }
    U256 ia = minv(acc, FR);
    for (size_t j = n; j-- > 0;) {
      U256 dj = mmul(ia, pre[j], FR);
      ia = mmul(ia, den[j], FR);
      L[j] = mmul(L[j], dj, FR);
    }
  }
  std::vector<U256> u(m, ZERO), v(m, ZERO), wv(m, ZERO);
  for (size_t j = 0; j < cs.ncons; j++) {
    for (const Entry& e : cs.A[j]) u[e.col] = madd(u[e.col], mmul(from_u64(e.coeff, FR), L[j], FR), FR);
    for (const Entry& e : cs.B[j]) v[e.col] = madd(v[e.col], mmul(from_u64(e.coeff, FR), L[j], FR), FR);
    for (const Entry& e : cs.C[j]) wv[e.col] = madd(wv[e.col], mmul(from_u64(e.coeff, FR), L[j], FR), FR);
  }
  for (size_t i = 0; i < 2; i++) u[i] = madd(u[i], L[cs.ncons + i], FR);
  U256 gi = minv(gamma, FR), di = minv(delta, FR);
  FixedBase<Fq> fb1(G1GEN);
  FixedBase<Fq2> fb2(g2gen());
  pk->alpha_g1 = fb1.mul(from_m(alpha, FR)).affine();
  pk->beta_g1 = fb1.mul(from_m(beta, FR)).affine();
  pk->delta_g1 = fb1.mul(from_m(delta, FR)).affine();
  pk->beta_g2 = fb2.mul(from_m(beta, FR)).affine();
  pk->delta_g2 = fb2.mul(from_m(delta, FR)).affine();
  pk->gamma_g2 = fb2.mul(from_m(gamma, FR)).affine();
  pk->a_query = fixed_mul_many(fb1, u);
  pk->b_g1_query = fixed_mul_many(fb1, v);
  pk->b_g2_query = fixed_mul_many(fb2, v);
  std::vector<U256> lq(m - 2), hq(n - 1), abc(2);
  for (size_t i = 0; i < m; i++) {
    U256 t = madd(madd(mmul(beta, u[i], FR), mmul(alpha, v[i], FR), FR), wv[i], FR);
    if (i < 2) abc[i] = mmul(t, gi, FR);
    else lq[i - 2] = mmul(t, di, FR);
  }
  U256 tp = mmul(zt, di, FR);
  for (size_t i = 0; i + 1 < n; i++) {
    hq[i] = tp;
    tp = mmul(tp, tau, FR);
  }
  pk->l_query = fixed_mul_many(fb1, lq);
  pk->h_query = fixed_mul_many(fb1, hq);
  pk->gamma_abc = fixed_mul_many(fb1, abc);
  return pk;
}

template <class F>
Aff<F> jadd_aff(const Aff<F>& a, const Aff<F>& b) {
  Jac<F> j = Jac<F>::identity();
  j.add_affine(a);
  j.add_affine(b);
  return j.affine();
}
template <class F>
Aff<F> aneg(const Aff<F>& a) { return {a.x, a.y.neg(), a.inf}; }

}  // namespace

// ================================================================== C interface (ctypes)
extern "C" {

int ob_threads() { return omp_get_max_threads(); }

void ob_fr_ntt(uint8_t* data, unsigned log_n, int inverse, int coset) {
  size_t n = (size_t)1 << log_n;
  std::vector<U256> a(n);
#pragma omp parallel for schedule(static)
  for (long i = 0; i < (long)n; i++) {
    U256 c;
    memcpy(c.w, data + 32 * (size_t)i, 32);
    a[i] = to_m(c, FR);
  }
  fr_ntt_full(a, log_n, inverse != 0, coset != 0);
#pragma omp parallel for schedule(static)
  for (long i = 0; i < (long)n; i++) {
    U256 c = from_m(a[i], FR);
    memcpy(data + 32 * (size_t)i, c.w, 32);
  }
}

// timing variant: conversion to/from Montgomery excluded; returns seconds of the transform alone
double ob_fr_ntt_timed(uint8_t* data, unsigned log_n, int inverse, int coset, int reps) {
  size_t n = (size_t)1 << log_n;
  std::vector<U256> a(n);
  for (size_t i = 0; i < n; i++) {
    U256 c;
    memcpy(c.w, data + 32 * i, 32);
    a[i] = to_m(c, FR);
  }
  double best = 1e30;
  for (int r = 0; r < reps; r++) {
    double t0 = omp_get_wtime();
    fr_ntt_full(a, log_n, inverse != 0, coset != 0);
    double t = omp_get_wtime() - t0;
    if (t < best) best = t;
  }
  for (size_t i = 0; i < n; i++) {
    U256 c = from_m(a[i], FR);
    memcpy(data + 32 * i, c.w, 32);
  }
  return best;
}

static std::vector<U256> load_scalars(const uint8_t* s, size_t n) {
  std::vector<U256> v(n);
  for (size_t i = 0; i < n; i++) memcpy(v[i].w, s + 32 * i, 32);
  return v;
}

double ob_msm_g1(const uint8_t* bases, const uint8_t* scalars, uint64_t n, uint8_t* out) {
  std::vector<Aff<Fq>> b(n);
#pragma omp parallel for schedule(static)
  for (long i = 0; i < (long)n; i++) b[i] = aff_get<Fq>(bases + 64 * (size_t)i);
  std::vector<U256> s = load_scalars(scalars, n);
  double t0 = omp_get_wtime();
  Jac<Fq> r = msm_arkworks(b.data(), s.data(), n);
  double t = omp_get_wtime() - t0;
  aff_put(r.affine(), out);
  return t;
}
double ob_msm_g2(const uint8_t* bases, const uint8_t* scalars, uint64_t n, uint8_t* out) {
  std::vector<Aff<Fq2>> b(n);
#pragma omp parallel for schedule(static)
  for (long i = 0; i < (long)n; i++) b[i] = aff_get<Fq2>(bases + 128 * (size_t)i);
  std::vector<U256> s = load_scalars(scalars, n);
  double t0 = omp_get_wtime();
  Jac<Fq2> r = msm_arkworks(b.data(), s.data(), n);
  double t = omp_get_wtime() - t0;
  aff_put(r.affine(), out);
  return t;
}

// P_i = k_i * G with k_i = xoshiro256**(seed) — the synthetic bases of the benchmark
void ob_synthetic_bases_g1(uint64_t n, uint64_t seed, uint8_t* out) {
  Xoshiro rng(seed);
  std::vector<U256> k(n);
  for (size_t i = 0; i < n; i++) k[i] = rng.fr();
  FixedBase<Fq> fb(G1GEN);
#pragma omp parallel for schedule(dynamic, 256)
  for (long i = 0; i < (long)n; i++) aff_put(fb.mul(k[i]).affine(), out + 64 * (size_t)i);
}
void ob_synthetic_bases_g2(uint64_t n, uint64_t seed, uint8_t* out) {
  Xoshiro rng(seed);
  std::vector<U256> k(n);
  for (size_t i = 0; i < n; i++) k[i] = rng.fr();
  FixedBase<Fq2> fb(g2gen());
#pragma omp parallel for schedule(dynamic, 64)
  for (long i = 0; i < (long)n; i++) aff_put(fb.mul(k[i]).affine(), out + 128 * (size_t)i);
}
void ob_random_scalars(uint64_t n, uint64_t seed, uint8_t* out) {
  Xoshiro rng(seed);
  for (size_t i = 0; i < n; i++) {
    U256 v = rng.fr();
    memcpy(out + 32 * i, v.w, 32);
  }
}

// ---- closed-form checkers for MSMs over the synthetic bases P_i = k_i * G:
//      sum_i s_i P_i = (sum_i s_i k_i mod r) * G, so a 2^24..2^26-point MSM is checked with one
//      O(n) dot product over Fr and ONE fixed-base multiplication.
void ob_set_threads(int n) {
  if (n > 0) omp_set_num_threads(n);
}
// out = sum_i a[i] * b[i] mod r   (canonical little-endian in and out)
void ob_fr_dot(const uint8_t* a, const uint8_t* b, uint64_t n, uint8_t* out) {
  const int T = omp_get_max_threads();
  std::vector<U256> part((size_t)T, U256{{0, 0, 0, 0}});
#pragma omp parallel
  {
    U256 acc = {{0, 0, 0, 0}};
#pragma omp for schedule(static)
    for (long i = 0; i < (long)n; i++) {
      U256 x, y;
      memcpy(x.w, a + 32 * (size_t)i, 32);
      memcpy(y.w, b + 32 * (size_t)i, 32);
      // mmul(x, y) = x*y/R; one more Montgomery product with R^2 after the sum restores x*y
      acc = madd(acc, mmul(x, y, FR), FR);
    }
    part[(size_t)omp_get_thread_num()] = acc;
  }
  U256 tot = {{0, 0, 0, 0}};
  for (int t = 0; t < T; t++) tot = madd(tot, part[(size_t)t], FR);
  tot = mmul(tot, FR.r2, FR);      // (sum x*y / R) * R^2 / R = sum x*y
  memcpy(out, tot.w, 32);
}
// out = sum_i a[i] mod r
void ob_fr_sum(const uint8_t* a, uint64_t n, uint8_t* out) {
  U256 tot = {{0, 0, 0, 0}};
  for (size_t i = 0; i < n; i++) {
    U256 x;
    memcpy(x.w, a + 32 * i, 32);
    tot = madd(tot, x, FR);
  }
  memcpy(out, tot.w, 32);
}
// out = k * G1 generator / k * G2 generator (canonical affine)
void ob_g1_mul_gen(const uint8_t* k, uint8_t* out) {
  U256 s;
  memcpy(s.w, k, 32);
  Jac<Fq> acc = Jac<Fq>::identity(), base = {G1GEN.x, G1GEN.y, Fq::one()};
  for (int i = 255; i >= 0; i--) {
    acc = acc.dbl();
    if ((s.w[i >> 6] >> (i & 63)) & 1) acc.add(base);
  }
  aff_put(acc.affine(), out);
}
void ob_g2_mul_gen(const uint8_t* k, uint8_t* out) {
  U256 s;
  memcpy(s.w, k, 32);
  Aff<Fq2> g = g2gen();
  Jac<Fq2> acc = Jac<Fq2>::identity(), base = {g.x, g.y, Fq2::one()};
  for (int i = 255; i >= 0; i--) {
    acc = acc.dbl();
    if ((s.w[i >> 6] >> (i & 63)) & 1) acc.add(base);
  }
  aff_put(acc.affine(), out);
}

void* ob_setup_synthetic(unsigned log_n, uint64_t seed) { return setup_synthetic(log_n, seed); }
void ob_pk_free(void* pk) { delete (ProvingKey*)pk; }

// vk layout of include/zkb.h: alpha_g1 | beta_g2 | gamma_g2 | delta_g2 | ic0 | ic1
void ob_pk_vk(void* pkv, uint8_t* out) {
  ProvingKey* pk = (ProvingKey*)pkv;
  aff_put(pk->alpha_g1, out);
  aff_put(pk->beta_g2, out + 64);
  aff_put(pk->gamma_g2, out + 192);
  aff_put(pk->delta_g2, out + 320);
  aff_put(pk->gamma_abc[0], out + 448);
  aff_put(pk->gamma_abc[1], out + 512);
}
// alpha_g1 | beta_g1 | delta_g1 (192 B) then beta_g2 | delta_g2 (256 B)
void ob_pk_fixed(void* pkv, uint8_t* out) {
  ProvingKey* pk = (ProvingKey*)pkv;
  aff_put(pk->alpha_g1, out);
  aff_put(pk->beta_g1, out + 64);
  aff_put(pk->delta_g1, out + 128);
  aff_put(pk->beta_g2, out + 192);
  aff_put(pk->delta_g2, out + 320);
}
// query export: which = 0 a, 1 b_g1, 2 b_g2, 3 l, 4 h ; returns the number of points
uint64_t ob_pk_query(void* pkv, int which, uint8_t* out) {
  ProvingKey* pk = (ProvingKey*)pkv;
  if (which == 2) {
    if (out) for (size_t i = 0; i < pk->b_g2_query.size(); i++) aff_put(pk->b_g2_query[i], out + 128 * i);
    return pk->b_g2_query.size();
  }
  const std::vector<Aff<Fq>>& q = which == 0 ? pk->a_query : which == 1 ? pk->b_g1_query : which == 3 ? pk->l_query : pk->h_query;
  if (out) for (size_t i = 0; i < q.size(); i++) aff_put(q[i], out + 64 * i);
  return q.size();
}

void ob_synthetic_witness(unsigned log_n, const uint8_t* w0, const uint8_t* w1, uint8_t* z_out) {
  size_t m = (size_t)1 << log_n;
  std::vector<U256> z(m, ZERO);
  U256 a, b;
  memcpy(a.w, w0, 32);
  memcpy(b.w, w1, 32);
  z[0] = FR.r1;
  z[2] = to_m(a, FR);
  z[3] = to_m(b, FR);
  for (size_t j = 0; j + 4 < m; j++) z[j + 4] = mmul(madd(z[j + 2], from_u64(j % 7 + 1, FR), FR), z[j + 3], FR);
  z[1] = mmul(z[m - 1], z[m - 2], FR);
  for (size_t i = 0; i < m; i++) {
    U256 c = from_m(z[i], FR);
    memcpy(z_out + 32 * i, c.w, 32);
  }
}

// returns seconds spent in the prover proper (witness map + 5 MSMs + assembly).
// proof_out: A 64 | B 128 | C 64.  inter_out (optional): msm_a|msm_b1|msm_b2|msm_l|msm_h (384 B).
// h_out (optional): (n-1) * 32 B.
double ob_groth16_prove(void* pkv, const uint8_t* witness, const uint8_t* r_bytes, const uint8_t* s_bytes,
                        uint8_t* proof_out, uint8_t* inter_out, uint8_t* h_out) {
  ProvingKey* pk = (ProvingKey*)pkv;
  const R1CS& cs = pk->cs;
  size_t n = pk->n, m = pk->m;
  std::vector<U256> zc = load_scalars(witness, m);   // canonical
  std::vector<U256> z(m);
  for (size_t i = 0; i < m; i++) z[i] = to_m(zc[i], FR);
  U256 r, s;
  memcpy(r.w, r_bytes, 32);
  memcpy(s.w, s_bytes, 32);
  double t0 = omp_get_wtime();
  std::vector<U256> a(n, ZERO), b(n, ZERO), c(n, ZERO);
#pragma omp parallel for schedule(static)
  for (long j = 0; j < (long)cs.ncons; j++) {
    a[j] = row_dot(cs.A[j], z);
    b[j] = row_dot(cs.B[j], z);
    c[j] = row_dot(cs.C[j], z);
  }
  a[cs.ncons] = z[0];
  a[cs.ncons + 1] = z[1];
  fr_ntt_full(a, pk->log_n, true, false);
  fr_ntt_full(b, pk->log_n, true, false);
  fr_ntt_full(c, pk->log_n, true, false);
  fr_ntt_full(a, pk->log_n, false, true);
  fr_ntt_full(b, pk->log_n, false, true);
  fr_ntt_full(c, pk->log_n, false, true);
  U256 zinv = minv(msub(mpow64(from_u64(5, FR), n, FR), FR.r1, FR), FR);
#pragma omp parallel for schedule(static)
  for (long i = 0; i < (long)n; i++) a[i] = mmul(msub(mmul(a[i], b[i], FR), c[i], FR), zinv, FR);
  fr_ntt_full(a, pk->log_n, true, true);
  std::vector<U256> h(n - 1);
  for (size_t i = 0; i + 1 < n; i++) h[i] = from_m(a[i], FR);
  Jac<Fq> ma = msm_arkworks(pk->a_query.data(), zc.data(), m);
  Jac<Fq> mb1 = msm_arkworks(pk->b_g1_query.data(), zc.data(), m);
  Jac<Fq2> mb2 = msm_arkworks(pk->b_g2_query.data(), zc.data(), m);
  Jac<Fq> ml = msm_arkworks(pk->l_query.data(), zc.data() + 2, m - 2);
  Jac<Fq> mh = msm_arkworks(pk->h_query.data(), h.data(), n - 1);
  // A = alpha + msm_a + r delta ; B = beta + msm_b + s delta ; C = msm_l + msm_h + sA + rB1 - rs delta
  Jac<Fq> A = ma;
  A.add_affine(pk->alpha_g1);
  A.add(scalar_mul(pk->delta_g1, r));
  Jac<Fq> B1 = mb1;
  B1.add_affine(pk->beta_g1);
  B1.add(scalar_mul(pk->delta_g1, s));
  Jac<Fq2> B2 = mb2;
  B2.add_affine(pk->beta_g2);
  B2.add(scalar_mul(pk->delta_g2, s));
  Aff<Fq> Aa = A.affine(), B1a = B1.affine();
  Jac<Fq> Cc = ml;
  Cc.add(mh);
  Cc.add(scalar_mul(Aa, s));
  Cc.add(scalar_mul(B1a, r));
  U256 rs = from_m(mmul(to_m(r, FR), to_m(s, FR), FR), FR);
  Cc.add_affine(aneg(scalar_mul(pk->delta_g1, rs).affine()));
  double t = omp_get_wtime() - t0;
  aff_put(Aa, proof_out);
  aff_put(B2.affine(), proof_out + 64);
  aff_put(Cc.affine(), proof_out + 192);
  if (inter_out) {
    aff_put(ma.affine(), inter_out);
    aff_put(mb1.affine(), inter_out + 64);
    aff_put(mb2.affine(), inter_out + 128);
    aff_put(ml.affine(), inter_out + 256);
    aff_put(mh.affine(), inter_out + 320);
  }
  if (h_out) memcpy(h_out, h.data(), (n - 1) * 32);
  return t;
}

}  // extern "C"
