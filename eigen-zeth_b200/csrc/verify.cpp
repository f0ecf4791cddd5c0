// verify.cpp — off-chain Groth16 verification (CPU): optimal-ate pairing on BN254.
//
// Mirrors what the consumer of our proofs does on-chain: EigenZkVM.verifyBatches / verifyTx
// (reference src/settlement/ethereum/interfaces/zkvm.rs:82-130, ABI in contracts/EigenZkVM.json)
// checks e(A,B) = e(alpha,beta) e(ic0 + x ic1, gamma) e(C,delta) with the bn256 precompiles.
// Running the same check before a proof leaves the prover is cheap insurance; it is host code by
// nature (a few milliseconds, O(1) in the circuit size) and is not a fallback for any GPU path.
//
// Tower: Fq2 = Fq[u]/(u^2+1), Fq6 = Fq2[v]/(v^3 - xi), xi = 9+u, Fq12 = Fq6[w]/(w^2 - v).
// D-type twist E': y^2 = x^3 + 3/xi, untwist (x, y) -> (x w^2, y w^3).
#include <cstdint>
#include <cstring>

#include "../../include/zkb.h"
#include "host_ff.hpp"

namespace {
using zkb::host::HFq;
using zkb::host::HFr;

struct F2 {
  HFq a, b;
  static F2 zero() { return {HFq::zero(), HFq::zero()}; }
  static F2 one() { return {HFq::one(), HFq::zero()}; }
  F2 operator+(const F2& o) const { return {a + o.a, b + o.b}; }
  F2 operator-(const F2& o) const { return {a - o.a, b - o.b}; }
  F2 operator*(const F2& o) const {
    HFq t0 = a * o.a, t1 = b * o.b;
    return {t0 - t1, (a + b) * (o.a + o.b) - t0 - t1};
  }
  F2 sqr() const { return {(a + b) * (a - b), (a * b).dbl()}; }
  F2 neg() const { return {a.neg(), b.neg()}; }
  F2 conj() const { return {a, b.neg()}; }
  F2 scale(const HFq& k) const { return {a * k, b * k}; }
  F2 inverse() const {
    HFq d = (a.sqr() + b.sqr()).inverse();
    return {a * d, (b * d).neg()};
  }
  F2 mul_xi() const {  // (a + b u)(9 + u) = 9a - b + (a + 9b) u
    HFq a2 = a.dbl(), a4 = a2.dbl(), a8 = a4.dbl(), b2 = b.dbl(), b4 = b2.dbl(), b8 = b4.dbl();
    return {a8 + a - b, b8 + b + a};
  }
  bool is_zero() const { return a.is_zero() && b.is_zero(); }
  bool operator==(const F2& o) const { return a == o.a && b == o.b; }
};

F2 f2_pow(F2 base, const uint64_t* e, int words) {
  F2 r = F2::one();
  for (int i = words * 64 - 1; i >= 0; i--) {
    r = r.sqr();
    if ((e[i >> 6] >> (i & 63)) & 1) r = r * base;
  }
  return r;
}

struct F6 {
  F2 c0, c1, c2;
  static F6 zero() { return {F2::zero(), F2::zero(), F2::zero()}; }
  static F6 one() { return {F2::one(), F2::zero(), F2::zero()}; }
  F6 operator+(const F6& o) const { return {c0 + o.c0, c1 + o.c1, c2 + o.c2}; }
  F6 operator-(const F6& o) const { return {c0 - o.c0, c1 - o.c1, c2 - o.c2}; }
  F6 neg() const { return {c0.neg(), c1.neg(), c2.neg()}; }
  F6 operator*(const F6& o) const {
    F2 t0 = c0 * o.c0, t1 = c1 * o.c1, t2 = c2 * o.c2;
    F2 r0 = t0 + ((c1 + c2) * (o.c1 + o.c2) - t1 - t2).mul_xi();
    F2 r1 = (c0 + c1) * (o.c0 + o.c1) - t0 - t1 + t2.mul_xi();
    F2 r2 = (c0 + c2) * (o.c0 + o.c2) - t0 - t2 + t1;
    return {r0, r1, r2};
  }
  F6 mul_v() const { return {c2.mul_xi(), c0, c1}; }
  F6 inverse() const {
    F2 t0 = c0.sqr() - (c1 * c2).mul_xi();
    F2 t1 = c2.sqr().mul_xi() - c0 * c1;
    F2 t2 = c1.sqr() - c0 * c2;
    F2 d = (c0 * t0 + (c2 * t1 + c1 * t2).mul_xi()).inverse();
    return {t0 * d, t1 * d, t2 * d};
  }
  bool operator==(const F6& o) const { return c0 == o.c0 && c1 == o.c1 && c2 == o.c2; }
};

struct F12 {
  F6 c0, c1;
  static F12 one() { return {F6::one(), F6::zero()}; }
  F12 operator*(const F12& o) const {
    F6 t0 = c0 * o.c0, t1 = c1 * o.c1;
    return {t0 + t1.mul_v(), (c0 + c1) * (o.c0 + o.c1) - t0 - t1};
  }
  F12 sqr() const { return *this * *this; }
  F12 conj() const { return {c0, c1.neg()}; }   // = x^{q^6}
  F12 inverse() const {
    F6 d = (c0 * c0 - (c1 * c1).mul_v()).inverse();
    return {c0 * d, (c1 * d).neg()};
  }
  bool operator==(const F12& o) const { return c0 == o.c0 && c1 == o.c1; }
};

struct G1 {
  HFq x, y;
  bool inf;
};
struct G2 {
  F2 x, y;
  bool inf;
};

const HFq B1 = HFq::from_u64(3);
F2 twist_b() {
  F2 xi{HFq::from_u64(9), HFq::one()};
  return F2{B1, HFq::zero()} * xi.inverse();
}

}  // namespace
namespace zkb {
// curve coefficient b in Montgomery form as 32-bit words: 8 words for G1 (b = 3), 16 for the twist
// (b' = 3 / (9 + u), c0 then c1) — used by the on-curve check of uploaded bases (msm_impl.cuh)
void host_curve_b(int is_g2, uint32_t* out) {
  if (!is_g2) {
    std::memcpy(out, B1.l, 32);
  } else {
    F2 b = twist_b();
    std::memcpy(out, b.a.l, 32);
    std::memcpy(out + 8, b.b.l, 32);
  }
}
}  // namespace zkb
namespace {

bool g1_on_curve(const G1& p) { return p.inf || p.y.sqr() == p.x.sqr() * p.x + B1; }
bool g2_on_curve(const G2& p) { return p.inf || p.y.sqr() == p.x.sqr() * p.x + twist_b(); }

G1 g1_add(const G1& p, const G1& q) {
  if (p.inf) return q;
  if (q.inf) return p;
  HFq lam;
  if (p.x == q.x) {
    if ((p.y + q.y).is_zero()) return {HFq::zero(), HFq::zero(), true};
    HFq xx = p.x.sqr();
    lam = (xx.dbl() + xx) * p.y.dbl().inverse();
  } else {
    lam = (q.y - p.y) * (q.x - p.x).inverse();
  }
  HFq x3 = lam.sqr() - p.x - q.x;
  return {x3, lam * (p.x - x3) - p.y, false};
}
G1 g1_mul(const G1& p, const uint64_t k[4]) {
  G1 r{HFq::zero(), HFq::zero(), true};
  for (int i = 255; i >= 0; i--) {
    r = g1_add(r, r);
    if ((k[i >> 6] >> (i & 63)) & 1) r = g1_add(r, p);
  }
  return r;
}
G2 g2_add(const G2& p, const G2& q) {
  if (p.inf) return q;
  if (q.inf) return p;
  F2 lam;
  if (p.x == q.x) {
    if ((p.y + q.y).is_zero()) return {F2::zero(), F2::zero(), true};
    F2 xx = p.x.sqr();
    lam = (xx + xx + xx) * (p.y + p.y).inverse();
  } else {
    lam = (q.y - p.y) * (q.x - p.x).inverse();
  }
  F2 x3 = lam.sqr() - p.x - q.x;
  return {x3, lam * (p.x - x3) - p.y, false};
}
// r * P == O  (subgroup membership on the twist; G1 has cofactor 1)
bool g2_in_subgroup(const G2& p) {
  if (p.inf) return true;
  G2 r{F2::zero(), F2::zero(), true};
  for (int i = 255; i >= 0; i--) {
    r = g2_add(r, r);
    if ((zkb::host::FrCfg::MOD[i >> 6] >> (i & 63)) & 1) r = g2_add(r, p);
  }
  return r.inf;
}

// line through T and S (tangent if equal) on the twist, evaluated at P; T <- T + S.
// l = yP - lam xP w + (lam xT - yT) w^3, w^3 = v w  ->  c0 = (yP,0,0), c1 = (-lam xP, lam xT - yT, 0)
F12 line(G2& T, const G2& S, const G1& P) {
  F2 lam;
  if (T.x == S.x && T.y == S.y) {
    F2 xx = T.x.sqr();
    lam = (xx + xx + xx) * (T.y + T.y).inverse();
  } else {
    lam = (S.y - T.y) * (S.x - T.x).inverse();
  }
  F2 x3 = lam.sqr() - T.x - S.x;
  F2 y3 = lam * (T.x - x3) - T.y;
  F12 l;
  l.c0 = {F2{P.y, HFq::zero()}, F2::zero(), F2::zero()};
  l.c1 = {lam.scale(P.x).neg(), lam * T.x - T.y, F2::zero()};
  T = {x3, y3, false};
  return l;
}

struct Frob {
  F2 gx, gy;   // xi^{(q-1)/3}, xi^{(q-1)/2}
  Frob() {
    F2 xi{HFq::from_u64(9), HFq::one()};
    // (q-1)/3 and (q-1)/2 as 256-bit integers
    uint64_t q1[4] = {zkb::host::FqCfg::MOD[0] - 1, zkb::host::FqCfg::MOD[1], zkb::host::FqCfg::MOD[2],
                      zkb::host::FqCfg::MOD[3]};
    uint64_t e3[4], e2[4];
    unsigned __int128 rem = 0;
    for (int i = 3; i >= 0; i--) {
      unsigned __int128 cur = (rem << 64) | q1[i];
      e3[i] = (uint64_t)(cur / 3);
      rem = cur % 3;
    }
    for (int i = 0; i < 4; i++) e2[i] = (q1[i] >> 1) | (i < 3 ? q1[i + 1] << 63 : 0);
    gx = f2_pow(xi, e3, 4);
    gy = f2_pow(xi, e2, 4);
  }
};
G2 frobenius(const G2& p, const Frob& f) { return {p.x.conj() * f.gx, p.y.conj() * f.gy, false}; }

const uint64_t ATE_LOOP[2] = {0x9d797039be763ba8ull, 0x1ull};   // 6x + 2 = 29793968203157093288

F12 miller(const G1& P, const G2& Q, const Frob& fr) {
  if (P.inf || Q.inf) return F12::one();
  F12 f = F12::one();
  G2 T = Q;
  for (int i = 63; i >= 0; i--) {   // bit 64 is the leading one
    f = f.sqr() * line(T, T, P);
    if ((ATE_LOOP[0] >> i) & 1) f = f * line(T, Q, P);
  }
  G2 Q1 = frobenius(Q, fr);
  G2 Q2 = frobenius(Q1, fr);
  Q2.y = Q2.y.neg();
  f = f * line(T, Q1, P);
  f = f * line(T, Q2, P);
  return f;
}

// f^{(q^12-1)/r}: easy part by conjugation/inversion and one generic power by q^2+1, hard part
// (q^4 - q^2 + 1)/r by plain square-and-multiply (761-bit exponent).
const uint64_t Q2P1[8] = {0x3b5458a2275d69b2ull, 0xa602072d09eac101ull, 0x4a50189c6d96cadcull,
                          0x04689e957a1242c8ull, 0x26edfa5c34c6b38dull, 0xb00b855116375606ull,
                          0x599a6f7c0348d21cull, 0x0925c4b8763cbf9cull};
const uint64_t HARD[12] = {0xe81bb482ccdf42b1ull, 0x5abf5cc4f49c36d4ull, 0xf1154e7e1da014fdull,
                           0xdcc7b44c87cdbacfull, 0xaaa441e3954bcf8aull, 0x6b887d56d5095f23ull,
                           0x79581e16f3fd90c6ull, 0x3b1b1355d189227dull, 0x4e529a5861876f6bull,
                           0x6c0eb522d5b12278ull, 0x331ec15183177fafull, 0x01baaa710b0759adull};

F12 f12_pow(const F12& b, const uint64_t* e, int words) {
  F12 r = F12::one();
  bool started = false;
  for (int i = words * 64 - 1; i >= 0; i--) {
    if (started) r = r.sqr();
    if ((e[i >> 6] >> (i & 63)) & 1) {
      r = started ? r * b : b;
      started = true;
    }
  }
  return r;
}
F12 final_exp(const F12& f) {
  F12 t = f.conj() * f.inverse();     // f^{q^6 - 1}
  t = f12_pow(t, Q2P1, 8);            // ^(q^2 + 1)
  return f12_pow(t, HARD, 12);        // ^((q^4 - q^2 + 1)/r)
}

bool read_fq(const uint8_t* b, HFq* out) { return HFq::from_bytes(b, out); }
bool read_g1(const uint8_t* b, G1* p) {
  bool z = true;
  for (int i = 0; i < 64; i++) z = z && !b[i];
  if (z) {
    *p = {HFq::zero(), HFq::zero(), true};
    return true;
  }
  p->inf = false;
  return read_fq(b, &p->x) && read_fq(b + 32, &p->y) && g1_on_curve(*p);
}
bool read_g2(const uint8_t* b, G2* p) {
  bool z = true;
  for (int i = 0; i < 128; i++) z = z && !b[i];
  if (z) {
    *p = {F2::zero(), F2::zero(), true};
    return true;
  }
  p->inf = false;
  return read_fq(b, &p->x.a) && read_fq(b + 32, &p->x.b) && read_fq(b + 64, &p->y.a) &&
         read_fq(b + 96, &p->y.b) && g2_on_curve(*p) && g2_in_subgroup(*p);
}

}  // namespace

extern "C" int zkb_groth16_verify(const uint8_t vk[ZKB_VK_BYTES], const uint8_t proof[ZKB_PROOF_BYTES],
                                  const uint8_t public_input[32]) {
  if (!vk || !proof || !public_input) return ZKB_EINVAL;
  G1 alpha, ic0, ic1, A, C;
  G2 beta, gamma, delta, B;
  if (!read_g1(vk, &alpha) || !read_g2(vk + 64, &beta) || !read_g2(vk + 192, &gamma) ||
      !read_g2(vk + 320, &delta) || !read_g1(vk + 448, &ic0) || !read_g1(vk + 512, &ic1))
    return ZKB_EINVAL;
  if (!read_g1(proof, &A) || !read_g2(proof + 64, &B) || !read_g1(proof + 192, &C)) return ZKB_EINVAL;
  uint64_t x[4];
  std::memcpy(x, public_input, 32);
  if (!HFr::lt_mod(x)) return ZKB_EINVAL;   // NewStateRootNotInsidePrime on-chain
  G1 acc = g1_add(ic0, g1_mul(ic1, x));
  static const Frob fr;
  auto negp = [](G1 p) { p.y = p.y.neg(); return p; };
  F12 f = miller(A, B, fr) * miller(negp(alpha), beta, fr) * miller(negp(acc), gamma, fr) *
          miller(negp(C), delta, fr);
  return final_exp(f) == F12::one() ? ZKB_OK : ZKB_EVERIFY;
}

extern "C" int zkb_proof_check(const uint8_t proof[ZKB_PROOF_BYTES], const uint8_t public_input[32],
                               uint32_t* reasons_opt) {
  if (!proof || !public_input) return ZKB_EINVAL;
  uint32_t bad = 0;
  auto all_zero = [](const uint8_t* b, int n) {
    bool z = true;
    for (int i = 0; i < n; i++) z = z && !b[i];
    return z;
  };
  auto check_g1 = [&](const uint8_t* b, uint32_t range_bit, uint32_t curve_bit) {
    if (all_zero(b, 64)) { bad |= ZKB_BAD_INFINITY; return; }
    G1 p{HFq::zero(), HFq::zero(), false};
    if (!read_fq(b, &p.x) || !read_fq(b + 32, &p.y)) { bad |= range_bit; return; }
    if (!g1_on_curve(p)) bad |= curve_bit;   // BN254 G1 has cofactor 1: on-curve = in the group
  };
  check_g1(proof, ZKB_BAD_A_RANGE, ZKB_BAD_A_CURVE);
  check_g1(proof + 192, ZKB_BAD_C_RANGE, ZKB_BAD_C_CURVE);
  const uint8_t* b = proof + 64;
  if (all_zero(b, 128)) {
    bad |= ZKB_BAD_INFINITY;
  } else {
    G2 p{F2::zero(), F2::zero(), false};
    if (!read_fq(b, &p.x.a) || !read_fq(b + 32, &p.x.b) || !read_fq(b + 64, &p.y.a) || !read_fq(b + 96, &p.y.b))
      bad |= ZKB_BAD_B_RANGE;
    else if (!g2_on_curve(p))
      bad |= ZKB_BAD_B_CURVE;
    else if (!g2_in_subgroup(p))
      bad |= ZKB_BAD_B_SUBGROUP;
  }
  uint64_t x[4];
  std::memcpy(x, public_input, 32);
  if (!HFr::lt_mod(x)) bad |= ZKB_BAD_INPUT_RANGE;
  if (reasons_opt) *reasons_opt = bad;
  return bad ? ZKB_EINVAL : ZKB_OK;
}
