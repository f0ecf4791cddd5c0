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
#include <mutex>
#include <new>
#include <thread>
#include <vector>

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
// k * P in Jacobian coordinates (one inversion at the end instead of one per ladder step)
G1 g1_mul(const G1& p, const uint64_t k[4]) {
  if (p.inf) return p;
  HFq X = HFq::one(), Y = HFq::one(), Z = HFq::zero();      // infinity
  for (int i = 255; i >= 0; i--) {
    if (!Z.is_zero()) {                                      // dbl-2009-l (a = 0)
      HFq a = X.sqr(), b = Y.sqr(), c = b.sqr();
      HFq d = (X + b).sqr() - a - c;
      d = d + d;
      HFq e = a + a + a, f = e.sqr();
      HFq x3 = f - d - d;
      HFq c8 = c + c;
      c8 = c8 + c8;
      c8 = c8 + c8;
      HFq z3 = Y * Z;
      Y = e * (d - x3) - c8;
      X = x3;
      Z = z3 + z3;
    }
    if ((k[i >> 6] >> (i & 63)) & 1) {
      if (Z.is_zero()) {
        X = p.x; Y = p.y; Z = HFq::one();
      } else {                                               // madd: affine p
        HFq z2 = Z.sqr();
        HFq u2 = p.x * z2, s2 = p.y * z2 * Z;
        HFq h = u2 - X, rr = s2 - Y;
        if (h.is_zero()) {
          if (!rr.is_zero()) { Z = HFq::zero(); continue; }  // P + (-P)
          // P + P: fall back to the affine doubling of p (cannot happen for k < r with P of order r,
          // kept for completeness)
          G1 d2 = g1_add(p, p);
          X = d2.x; Y = d2.y; Z = d2.inf ? HFq::zero() : HFq::one();
          continue;
        }
        HFq h2 = h.sqr(), h3 = h2 * h, v = X * h2;
        HFq x3 = rr.sqr() - h3 - v - v;
        Y = rr * (v - x3) - Y * h3;
        X = x3;
        Z = Z * h;
      }
    }
  }
  if (Z.is_zero()) return {HFq::zero(), HFq::zero(), true};
  HFq zi = Z.inverse(), zi2 = zi.sqr();
  return {X * zi2, Y * zi2 * zi, false};
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
// r * P == O  (subgroup membership on the twist; G1 has cofactor 1).  Jacobian coordinates: no field
// inversion inside the 254-step ladder (the affine ladder spent ~400 Fq multiplications per step on it).
struct G2J {
  F2 x, y, z;   // affine (x / z^2, y / z^3); z == 0: infinity
};
G2J g2j_dbl(const G2J& p) {
  if (p.z.is_zero()) return p;
  F2 a = p.x.sqr(), b = p.y.sqr(), c = b.sqr();
  F2 d = (p.x + b).sqr() - a - c;
  d = d + d;
  F2 e = a + a + a, f = e.sqr();
  F2 x3 = f - d - d;
  F2 c8 = c + c;
  c8 = c8 + c8;
  c8 = c8 + c8;
  F2 y3 = e * (d - x3) - c8;
  F2 z3 = p.y * p.z;
  return {x3, y3, z3 + z3};
}
G2J g2j_add_affine(const G2J& p, const G2& q) {     // q affine, not infinity
  if (p.z.is_zero()) return {q.x, q.y, F2::one()};
  F2 z2 = p.z.sqr();
  F2 u2 = q.x * z2, s2 = q.y * z2 * p.z;
  F2 h = u2 - p.x, rr = s2 - p.y;
  if (h.is_zero()) {
    if (rr.is_zero()) return g2j_dbl(p);
    return {F2::one(), F2::one(), F2::zero()};
  }
  F2 h2 = h.sqr(), h3 = h2 * h, v = p.x * h2;
  F2 x3 = rr.sqr() - h3 - v - v;
  F2 y3 = rr * (v - x3) - p.y * h3;
  return {x3, y3, p.z * h};
}
bool g2_in_subgroup(const G2& p) {
  if (p.inf) return true;
  G2J r{F2::one(), F2::one(), F2::zero()};
  for (int i = 255; i >= 0; i--) {
    r = g2j_dbl(r);
    if ((zkb::host::FrCfg::MOD[i >> 6] >> (i & 63)) & 1) r = g2j_add_affine(r, p);
  }
  return r.z.is_zero();
}

// Lines of the optimal-ate Miller loop, for up to four pairs at once.  Step j of pair k needs the
// slope lam = num / den on the twist (tangent: 3 x^2 / 2 y, chord: (yS - yT) / (xS - xT)); the
// denominators of all pairs are inverted together (Montgomery's trick on the Fq norms: ONE field
// inversion per step instead of one per pair), then each pair finishes its line
//   l = yP - lam xP w + (lam xT - yT) w^3,  w^3 = v w  ->  c0 = (yP,0,0), c1 = (-lam xP, lam xT - yT, 0)
// and moves T <- T + S.
constexpr int kMaxPairs = 4;
void batch_inverse_f2(F2* v, int n) {          // v[i] <- 1 / v[i]; all non-zero
  HFq norm[kMaxPairs], pre[kMaxPairs];
  HFq acc = HFq::one();
  for (int i = 0; i < n; i++) {
    norm[i] = v[i].a.sqr() + v[i].b.sqr();
    pre[i] = acc;
    acc = acc * norm[i];
  }
  HFq inv = acc.inverse();
  for (int i = n - 1; i >= 0; i--) {
    HFq ni = inv * pre[i];
    inv = inv * norm[i];
    v[i] = {v[i].a * ni, (v[i].b * ni).neg()};
  }
}
// one step for every pair: T[k] <- T[k] + S[k] (S == nullptr: doubling); f <- f * prod_k line_k
void line_step(F12& f, G2* T, const G2* S, const G1* P, int n) {
  F2 num[kMaxPairs], den[kMaxPairs];
  for (int k = 0; k < n; k++) {
    if (!S || (T[k].x == S[k].x && T[k].y == S[k].y)) {
      F2 xx = T[k].x.sqr();
      num[k] = xx + xx + xx;
      den[k] = T[k].y + T[k].y;
    } else {
      num[k] = S[k].y - T[k].y;
      den[k] = S[k].x - T[k].x;
    }
  }
  batch_inverse_f2(den, n);
  for (int k = 0; k < n; k++) {
    const G2& Sk = S ? S[k] : T[k];
    F2 lam = num[k] * den[k];
    F2 x3 = lam.sqr() - T[k].x - Sk.x;
    F2 y3 = lam * (T[k].x - x3) - T[k].y;
    F12 l;
    l.c0 = {F2{P[k].y, HFq::zero()}, F2::zero(), F2::zero()};
    l.c1 = {lam.scale(P[k].x).neg(), lam * T[k].x - T[k].y, F2::zero()};
    T[k] = {x3, y3, false};
    f = f * l;
  }
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

// prod_k f_{6x+2, Q_k}(P_k) with ONE shared accumulator: one F12 squaring per loop step whatever the
// number of pairs.  Pairs with a point at infinity contribute 1 and are dropped by the caller.
F12 multi_miller(const G1* P, const G2* Q, int n, const Frob& fr) {
  F12 f = F12::one();
  if (n == 0) return f;
  G2 T[kMaxPairs], Q1[kMaxPairs], Q2[kMaxPairs];
  for (int k = 0; k < n; k++) T[k] = Q[k];
  for (int i = 63; i >= 0; i--) {   // bit 64 is the leading one
    f = f.sqr();
    line_step(f, T, nullptr, P, n);
    if ((ATE_LOOP[0] >> i) & 1) line_step(f, T, Q, P, n);
  }
  for (int k = 0; k < n; k++) {
    Q1[k] = frobenius(Q[k], fr);
    Q2[k] = frobenius(Q1[k], fr);
    Q2[k].y = Q2[k].y.neg();
  }
  line_step(f, T, Q1, P, n);
  line_step(f, T, Q2, P, n);
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
// reference form: plain square-and-multiply on both exponents (kept as the checker of the fast one)
F12 final_exp_slow(const F12& f) {
  F12 t = f.conj() * f.inverse();     // f^{q^6 - 1}
  t = f12_pow(t, Q2P1, 8);            // ^(q^2 + 1)
  return f12_pow(t, HARD, 12);        // ^((q^4 - q^2 + 1)/r)
}

// Frobenius on Fq12 = Fq2[w]/(w^6 - xi): (sum c_i w^i)^(q^k) = sum conj^k(c_i) g_{k,i} w^i with
// g_{k,i} = xi^(i (q^k - 1)/6).  In the tower basis w^0..w^5 = c0.c0, c1.c0, c0.c1, c1.c1, c0.c2, c1.c2.
struct Frob12 {
  F2 g[4][6];     // g[k][i], k = 1, 2, 3
  Frob12() {
    F2 xi{HFq::from_u64(9), HFq::one()};
    uint64_t q1[4] = {zkb::host::FqCfg::MOD[0] - 1, zkb::host::FqCfg::MOD[1], zkb::host::FqCfg::MOD[2],
                      zkb::host::FqCfg::MOD[3]};
    uint64_t e6[4];
    unsigned __int128 rem = 0;
    for (int i = 3; i >= 0; i--) {
      unsigned __int128 cur = (rem << 64) | q1[i];
      e6[i] = (uint64_t)(cur / 6);
      rem = cur % 6;
    }
    F2 g1 = f2_pow(xi, e6, 4);        // xi^((q-1)/6)
    F2 p = F2::one();
    for (int i = 0; i < 6; i++) {
      g[1][i] = p;
      g[2][i] = p * p.conj();         // ^(q+1): the norm, in Fq
      g[3][i] = g[2][i] * p;          // ^(q^2+q+1)
      p = p * g1;
    }
  }
  F12 apply(const F12& f, int k) const {
    const F2* c[6] = {&f.c0.c0, &f.c1.c0, &f.c0.c1, &f.c1.c1, &f.c0.c2, &f.c1.c2};
    F2 r[6];
    for (int i = 0; i < 6; i++) r[i] = ((k & 1) ? c[i]->conj() : *c[i]) * g[k][i];
    return {{r[0], r[2], r[4]}, {r[1], r[3], r[5]}};
  }
};

F12 f12_pow_x(const F12& b) {         // ^x, x = 4965661367192848881 (the BN parameter)
  static const uint64_t X[1] = {4965661367192848881ull};
  return f12_pow(b, X, 1);
}

// f^{(q^12-1)/r}: easy part with one inversion and one Frobenius; hard part (q^4 - q^2 + 1)/r by the
// Devegili / Scott decomposition in base q with three powers by x (63 bits each) — ~260 Fq12
// operations instead of ~1900 for plain square-and-multiply.  In the cyclotomic subgroup the
// inverse is the conjugate.  tests/test_abi_host.py compares it with final_exp_slow on pairing values.
F12 final_exp(const F12& f) {
  static const Frob12 fb;
  F12 t1 = f.conj() * f.inverse();                  // f^(q^6 - 1)
  t1 = fb.apply(t1, 2) * t1;                        // ^(q^2 + 1)
  F12 fp = fb.apply(t1, 1), fp2 = fb.apply(t1, 2), fp3 = fb.apply(fp2, 1);
  F12 fu = f12_pow_x(t1), fu2 = f12_pow_x(fu), fu3 = f12_pow_x(fu2);
  F12 y3 = fb.apply(fu, 1).conj();
  F12 fu2p = fb.apply(fu2, 1), fu3p = fb.apply(fu3, 1);
  F12 y2 = fb.apply(fu2, 2);
  F12 y0 = fp * fp2 * fp3;
  F12 y1 = t1.conj();
  F12 y5 = fu2.conj();
  F12 y4 = (fu * fu2p).conj();
  F12 y6 = (fu3 * fu3p).conj();
  F12 t0 = y6.sqr() * y4 * y5;
  F12 s1 = y3 * y5 * t0;
  t0 = t0 * y2;
  s1 = (s1.sqr() * t0).sqr();
  t0 = s1 * y1;
  s1 = s1 * y0;
  return t0.sqr() * s1;
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

// e(A,B) e(-alpha,beta) e(-acc,gamma) e(-C,delta) == 1: one shared Miller loop, one final exponentiation
int pairing_check(const G1& A, const G2& B, const G1& C, const G1& alpha, const G2& beta, const G2& gamma,
                  const G2& delta, const G1& acc) {
  static const Frob fr;
  auto negp = [](G1 p) { p.y = p.y.neg(); return p; };
  const G1 Ps[4] = {A, negp(alpha), negp(acc), negp(C)};
  const G2 Qs[4] = {B, beta, gamma, delta};
  G1 P[4];
  G2 Q[4];
  int n = 0;
  for (int k = 0; k < 4; k++)
    if (!Ps[k].inf && !Qs[k].inf) {
      P[n] = Ps[k];
      Q[n++] = Qs[k];
    }
  return final_exp(multi_miller(P, Q, n, fr)) == F12::one() ? ZKB_OK : ZKB_EVERIFY;
}

}  // namespace

extern "C" int zkb_groth16_verify(const uint8_t vk[ZKB_VK_BYTES], const uint8_t proof[ZKB_PROOF_BYTES],
                                  const uint8_t public_input[32]) {
  if (!vk || !proof || !public_input) return ZKB_EINVAL;
  G1 alpha, ic0, ic1, A, C;
  G2 beta, gamma, delta, B;
  // a verifier checks many proofs against one key: the key's own validation (on-curve, three G2
  // subgroup checks) is remembered for the last key seen
  {
    static std::mutex mu;
    static uint8_t last_vk[ZKB_VK_BYTES];
    static bool have = false;
    static G1 c_alpha, c_ic0, c_ic1;
    static G2 c_beta, c_gamma, c_delta;
    std::lock_guard<std::mutex> lk(mu);
    if (!have || std::memcmp(last_vk, vk, ZKB_VK_BYTES) != 0) {
      have = false;
      if (!read_g1(vk, &c_alpha) || !read_g2(vk + 64, &c_beta) || !read_g2(vk + 192, &c_gamma) ||
          !read_g2(vk + 320, &c_delta) || !read_g1(vk + 448, &c_ic0) || !read_g1(vk + 512, &c_ic1))
        return ZKB_EINVAL;
      std::memcpy(last_vk, vk, ZKB_VK_BYTES);
      have = true;
    }
    alpha = c_alpha; ic0 = c_ic0; ic1 = c_ic1;
    beta = c_beta; gamma = c_gamma; delta = c_delta;
  }
  if (!read_g1(proof, &A) || !read_g2(proof + 64, &B) || !read_g1(proof + 192, &C)) return ZKB_EINVAL;
  uint64_t x[4];
  std::memcpy(x, public_input, 32);
  if (!HFr::lt_mod(x)) return ZKB_EINVAL;   // NewStateRootNotInsidePrime on-chain
  return pairing_check(A, B, C, alpha, beta, gamma, delta, g1_add(ic0, g1_mul(ic1, x)));
}

// any number of public inputs (keys of circuits other than the final-proof one): the key is
// validated on every call
extern "C" int zkb_groth16_verify_ex(const uint8_t vk_fixed[448], const uint8_t* ic, uint32_t n_public,
                                     const uint8_t proof[ZKB_PROOF_BYTES], const uint8_t* public_inputs) {
  if (!vk_fixed || !ic || !proof || (n_public && !public_inputs)) return ZKB_EINVAL;
  G1 alpha, A, C, acc;
  G2 beta, gamma, delta, B;
  if (!read_g1(vk_fixed, &alpha) || !read_g2(vk_fixed + 64, &beta) || !read_g2(vk_fixed + 192, &gamma) ||
      !read_g2(vk_fixed + 320, &delta) || !read_g1(ic, &acc))
    return ZKB_EINVAL;
  if (!read_g1(proof, &A) || !read_g2(proof + 64, &B) || !read_g1(proof + 192, &C)) return ZKB_EINVAL;
  for (uint32_t i = 0; i < n_public; i++) {
    G1 p;
    uint64_t x[4];
    std::memcpy(x, public_inputs + 32ull * i, 32);
    if (!read_g1(ic + 64ull * (i + 1), &p) || !HFr::lt_mod(x)) return ZKB_EINVAL;
    acc = g1_add(acc, g1_mul(p, x));
  }
  return pairing_check(A, B, C, alpha, beta, gamma, delta, acc);
}

// ---------------------------------------------------------------- proof assembly (K15) on the host
// A = alpha + msm_a + r delta1,  B = beta2 + msm_b2 + s delta2,  B1 = beta1 + msm_b1 + s delta1,
// C = msm_l + msm_h + s A + r B1 - r s delta1          
// Six 254-bit scalar multiplications and a few additions: O(1) in the circuit size and strictly
// serial (a ladder is 254 dependent doublings).  One GPU lane needs ~3 us per doubling (the
// IMAD.WIDE pipe gives a lone warp one instruction per 4 cycles), i.e. 1 ms per G1 ladder and 3 ms
// for the G2 one — measured 8 ms per proof with full-size r, s; a CPU core runs the same ladder
// in 0.15 ms.  So the five MSM sums (384 bytes) come back to the host, which already receives the
// proof anyway, and the ladders run here, the independent ones on separate threads.
namespace {
struct G1J {
  HFq x, y, z;
};
G1J g1j_dbl(const G1J& p) {
  if (p.z.is_zero()) return p;
  HFq a = p.x.sqr(), b = p.y.sqr(), c = b.sqr();
  HFq d = (p.x + b).sqr() - a - c;
  d = d + d;
  HFq e = a + a + a, f = e.sqr();
  HFq x3 = f - d - d;
  HFq c8 = c + c;
  c8 = c8 + c8;
  c8 = c8 + c8;
  HFq z3 = p.y * p.z;
  return {x3, e * (d - x3) - c8, z3 + z3};
}
G1J g1j_add_affine(const G1J& p, const G1& q) {
  if (q.inf) return p;
  if (p.z.is_zero()) return {q.x, q.y, HFq::one()};
  HFq z2 = p.z.sqr();
  HFq u2 = q.x * z2, s2 = q.y * z2 * p.z;
  HFq h = u2 - p.x, rr = s2 - p.y;
  if (h.is_zero()) {
    if (rr.is_zero()) return g1j_dbl(p);
    return {HFq::one(), HFq::one(), HFq::zero()};
  }
  HFq h2 = h.sqr(), h3 = h2 * h, v = p.x * h2;
  HFq x3 = rr.sqr() - h3 - v - v;
  return {x3, rr * (v - x3) - p.y * h3, p.z * h};
}
G1 g1j_affine(const G1J& p) {
  if (p.z.is_zero()) return {HFq::zero(), HFq::zero(), true};
  HFq zi = p.z.inverse(), zi2 = zi.sqr();
  return {p.x * zi2, p.y * zi2 * zi, false};
}
G1 g1_ladder(const G1& p, const uint64_t k[4]) {
  if (p.inf) return p;
  G1J r{HFq::one(), HFq::one(), HFq::zero()};
  for (int i = 255; i >= 0; i--) {
    r = g1j_dbl(r);
    if ((k[i >> 6] >> (i & 63)) & 1) r = g1j_add_affine(r, p);
  }
  return g1j_affine(r);
}
G2 g2j_affine(const G2J& p) {
  if (p.z.is_zero()) return {F2::zero(), F2::zero(), true};
  F2 zi = p.z.inverse(), zi2 = zi.sqr();
  return {p.x * zi2, p.y * zi2 * zi, false};
}
G2 g2_ladder(const G2& p, const uint64_t k[4]) {
  if (p.inf) return p;
  G2J r{F2::one(), F2::one(), F2::zero()};
  for (int i = 255; i >= 0; i--) {
    r = g2j_dbl(r);
    if ((k[i >> 6] >> (i & 63)) & 1) r = g2j_add_affine(r, p);
  }
  return g2j_affine(r);
}
// full Jacobian addition (both operands projective) and fixed-base window tables: the ladders on
// delta_g1 / delta_g2 are the same for every proof of a key, so each key carries, per 8-bit window w,
// the multiples d * 2^(8w) * delta (d = 1..255) and k * delta is 32 table additions.
template <class J, class F, class Dbl>
J jac_add(const J& p, const J& q, Dbl dbl) {
  if (p.z.is_zero()) return q;
  if (q.z.is_zero()) return p;
  F z1z1 = p.z.sqr(), z2z2 = q.z.sqr();
  F u1 = p.x * z2z2, u2 = q.x * z1z1;
  F s1 = p.y * q.z * z2z2, s2 = q.y * p.z * z1z1;
  F h = u2 - u1, rr = s2 - s1;
  if (h.is_zero()) {
    if (rr.is_zero()) return dbl(p);
    return {F::one(), F::one(), F::zero()};
  }
  F h2 = h.sqr(), h3 = h2 * h, v = u1 * h2;
  F x3 = rr.sqr() - h3 - v - v;
  return {x3, rr * (v - x3) - s1 * h3, p.z * q.z * h};
}

template <class J>
struct FixedTable {
  std::vector<J> t;      // [window][digit - 1], 32 x 255
};
struct KeyTables {
  FixedTable<G1J> d1;
  FixedTable<G2J> d2;
};
template <class J, class F, class Dbl>
void build_table(FixedTable<J>& tab, J base, Dbl dbl) {
  tab.t.resize(32 * 255);
  for (int w = 0; w < 32; w++) {
    J acc = base;
    for (int d = 1; d <= 255; d++) {
      tab.t[size_t(w) * 255 + d - 1] = acc;
      acc = jac_add<J, F>(acc, base, dbl);
    }
    base = acc;          // 256 * previous base
  }
}
template <class J, class F, class Dbl>
J table_mul(const FixedTable<J>& tab, const uint64_t k[4], Dbl dbl) {
  J r{F::one(), F::one(), F::zero()};
  for (int w = 0; w < 32; w++) {
    unsigned d = unsigned(k[w >> 3] >> ((w & 7) * 8)) & 255u;
    if (d) r = jac_add<J, F>(r, tab.t[size_t(w) * 255 + d - 1], dbl);
  }
  return r;
}

void put_g1(const G1& p, uint8_t* out) {
  if (p.inf) {
    std::memset(out, 0, 64);
    return;
  }
  p.x.to_bytes(out);
  p.y.to_bytes(out + 32);
}
void put_g2(const G2& p, uint8_t* out) {
  if (p.inf) {
    std::memset(out, 0, 128);
    return;
  }
  p.x.a.to_bytes(out);
  p.x.b.to_bytes(out + 32);
  p.y.a.to_bytes(out + 64);
  p.y.b.to_bytes(out + 96);
}
// points that come from our own GPU (MSM sums, key constants): canonical and on the curve by
// construction; only the encoding is decoded here
bool get_g1(const uint8_t* b, G1* p) {
  bool z = true;
  for (int i = 0; i < 64; i++) z = z && !b[i];
  if (z) {
    *p = {HFq::zero(), HFq::zero(), true};
    return true;
  }
  p->inf = false;
  return read_fq(b, &p->x) && read_fq(b + 32, &p->y);
}
bool get_g2(const uint8_t* b, G2* p) {
  bool z = true;
  for (int i = 0; i < 128; i++) z = z && !b[i];
  if (z) {
    *p = {F2::zero(), F2::zero(), true};
    return true;
  }
  p->inf = false;
  return read_fq(b, &p->x.a) && read_fq(b + 32, &p->x.b) && read_fq(b + 64, &p->y.a) && read_fq(b + 96, &p->y.b);
}
}  // namespace

namespace zkb {
// per-key window tables of delta_g1 / delta_g2 (1.5 MB, built once in ~20 ms); opaque to groth16.cu
void* host_assemble_tables_create(const uint8_t fixed[448]) {
  G1 delta1;
  G2 delta2;
  if (!get_g1(fixed + 128, &delta1) || !get_g2(fixed + 320, &delta2) || delta1.inf || delta2.inf) return nullptr;
  KeyTables* kt = new (std::nothrow) KeyTables();
  if (!kt) return nullptr;
  try {
    build_table<G1J, HFq>(kt->d1, G1J{delta1.x, delta1.y, HFq::one()}, g1j_dbl);
    build_table<G2J, F2>(kt->d2, G2J{delta2.x, delta2.y, F2::one()}, g2j_dbl);
  } catch (...) {
    delete kt;
    return nullptr;
  }
  return kt;
}
void host_assemble_tables_free(void* t) { delete static_cast<KeyTables*>(t); }

// fixed: alpha_g1 64 | beta_g1 64 | delta_g1 64 | beta_g2 128 | delta_g2 128 (canonical);
// sums: msm_a 64 | msm_b1 64 | msm_b2 128 | msm_l 64 | msm_h 64; r, s canonical < r;
// tables: from host_assemble_tables_create for THIS key's delta, or null (plain ladders)
bool host_groth16_assemble(const uint8_t fixed[448], const uint8_t sums[ZKB_INTERMEDIATE_BYTES], const uint8_t r[32],
                           const uint8_t s[32], uint8_t proof_out[ZKB_PROOF_BYTES], const void* tables) {
  G1 alpha, beta1, delta1, ma, mb1, ml, mh;
  G2 beta2, delta2, mb2;
  if (!get_g1(fixed, &alpha) || !get_g1(fixed + 64, &beta1) || !get_g1(fixed + 128, &delta1) ||
      !get_g2(fixed + 192, &beta2) || !get_g2(fixed + 320, &delta2) || !get_g1(sums, &ma) || !get_g1(sums + 64, &mb1) ||
      !get_g2(sums + 128, &mb2) || !get_g1(sums + 256, &ml) || !get_g1(sums + 320, &mh))
    return false;
  uint64_t rk[4], sk[4];
  std::memcpy(rk, r, 32);
  std::memcpy(sk, s, 32);
  if (!HFr::lt_mod(rk) || !HFr::lt_mod(sk)) return false;
  HFr rm{{rk[0], rk[1], rk[2], rk[3]}}, sm{{sk[0], sk[1], sk[2], sk[3]}};
  HFr rs = (rm.to_mont() * sm.to_mont()).from_mont();
  const KeyTables* kt = static_cast<const KeyTables*>(tables);
  // round 1: the multiples of delta (table additions with a key's tables, ladders without)
  G1 rd1, sd1, rsd1;
  G2 sd2;
  if (kt) {
    rd1 = g1j_affine(table_mul<G1J, HFq>(kt->d1, rk, g1j_dbl));
    sd1 = g1j_affine(table_mul<G1J, HFq>(kt->d1, sk, g1j_dbl));
    rsd1 = g1j_affine(table_mul<G1J, HFq>(kt->d1, rs.l, g1j_dbl));
    sd2 = g2j_affine(table_mul<G2J, F2>(kt->d2, sk, g2j_dbl));
  } else {
    std::thread t_g2([&] { sd2 = g2_ladder(delta2, sk); });
    std::thread t_sd1([&] { sd1 = g1_ladder(delta1, sk); });
    std::thread t_rs([&] { rsd1 = g1_ladder(delta1, rs.l); });
    rd1 = g1_ladder(delta1, rk);
    t_g2.join();
    t_sd1.join();
    t_rs.join();
  }
  G1 A = g1_add(g1_add(alpha, ma), rd1);
  G1 B1 = g1_add(g1_add(beta1, mb1), sd1);
  // round 2: the two variable-base ladders s A and r B1, side by side
  G1 sA, rB1;
  std::thread t_sa([&] { sA = g1_ladder(A, sk); });
  rB1 = g1_ladder(B1, rk);
  t_sa.join();
  rsd1.y = rsd1.y.neg();
  G1 C = g1_add(g1_add(g1_add(g1_add(ml, mh), sA), rB1), rsd1);
  G2 B = g2_add(g2_add(beta2, mb2), sd2);
  put_g1(A, proof_out);
  put_g2(B, proof_out + 64);
  put_g1(C, proof_out + 192);
  return true;
}
}  // namespace zkb

extern "C" int zkb_groth16_assemble_host(const uint8_t fixed[448], const uint8_t sums[ZKB_INTERMEDIATE_BYTES],
                                         const uint8_t r[32], const uint8_t s[32], uint8_t proof_out[ZKB_PROOF_BYTES]) {
  if (!fixed || !sums || !r || !s || !proof_out) return ZKB_EINVAL;
  return zkb::host_groth16_assemble(fixed, sums, r, s, proof_out, nullptr) ? ZKB_OK : ZKB_EINVAL;
}

// test hook: the fixed-base-table form of the assembly (what the prove calls use) against the ladder form
extern "C" int zkb_selftest_assemble_tables(const uint8_t fixed[448], const uint8_t sums[ZKB_INTERMEDIATE_BYTES],
                                            const uint8_t r[32], const uint8_t s[32]) {
  if (!fixed || !sums || !r || !s) return ZKB_EINVAL;
  void* t = zkb::host_assemble_tables_create(fixed);
  if (!t) return ZKB_ENOMEM;
  uint8_t a[ZKB_PROOF_BYTES], b[ZKB_PROOF_BYTES];
  const bool ok = zkb::host_groth16_assemble(fixed, sums, r, s, a, t) && zkb::host_groth16_assemble(fixed, sums, r, s, b, nullptr);
  zkb::host_assemble_tables_free(t);
  if (!ok) return ZKB_EINVAL;
  return std::memcmp(a, b, ZKB_PROOF_BYTES) == 0 ? ZKB_OK : ZKB_EVERIFY;
}

// test hook (tests/test_abi_host.py): 0 when the decomposed final exponentiation agrees with plain
// square-and-multiply on e(k1 G1, G2-point)-style Miller values built from the given proof points
extern "C" int zkb_selftest_final_exp(const uint8_t proof[ZKB_PROOF_BYTES]) {
  G1 A, C;
  G2 B;
  if (!proof || !read_g1(proof, &A) || !read_g2(proof + 64, &B) || !read_g1(proof + 192, &C)) return ZKB_EINVAL;
  if (A.inf || B.inf || C.inf) return ZKB_EINVAL;
  static const Frob fr;
  const G1 P[2] = {A, C};
  const G2 Q[2] = {B, B};
  for (int n = 1; n <= 2; n++) {
    F12 m = multi_miller(P, Q, n, fr);
    if (!(final_exp(m) == final_exp_slow(m))) return ZKB_EVERIFY;
  }
  return ZKB_OK;
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
