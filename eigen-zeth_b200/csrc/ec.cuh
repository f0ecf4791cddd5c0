// ec.cuh — short-Weierstrass (a = 0) point arithmetic in XYZZ coordinates, generic over the
// coordinate field F (Fq for BN254 G1, Fq2 for the G2 twist).
//
// No reference counterpart: this is the curve layer under the
// G1/G2 MSMs of the external prover behind reference proto/prover/v1/prover.proto:130-148.
//
// XYZZ: x = X/ZZ, y = Y/ZZZ with ZZ^3 = ZZZ^2; the point at infinity is ZZ == 0.
// Affine points are (x, y) in Montgomery form; the all-zero encoding is the point at infinity
// ((0,0) is not on y^2 = x^3 + b for b != 0).
#pragma once
#include "ff.cuh"

namespace zkb {
#ifdef __CUDACC__

template <class F>
struct Affine {
  F x, y;
  __device__ __forceinline__ bool is_inf() const { return x.raw_is_zero() && y.raw_is_zero(); }
  __device__ __forceinline__ static Affine load_nc(const void* p) {
    const char* c = reinterpret_cast<const char*>(p);
    return {F::load_nc(c), F::load_nc(c + F::kBytes)};
  }
  __device__ __forceinline__ static Affine load(const void* p) {
    const char* c = reinterpret_cast<const char*>(p);
    return {F::load(c), F::load(c + F::kBytes)};
  }
  __device__ __forceinline__ void store(void* p) const {
    char* c = reinterpret_cast<char*>(p);
    x.store(c);
    y.store(c + F::kBytes);
  }
};

template <class F>
struct XYZZ {
  F x, y, zz, zzz;

  __device__ __forceinline__ static XYZZ infinity() {
    return {F::zero(), F::zero(), F::zero(), F::zero()};
  }
  __device__ __forceinline__ bool is_inf() const { return zz.is_zero(); }
  __device__ __forceinline__ static XYZZ from_affine(const Affine<F>& p) {
    if (p.is_inf()) return infinity();
    return {p.x, p.y, F::one(), F::one()};
  }
  __device__ __forceinline__ static XYZZ load(const void* p) {
    const char* c = reinterpret_cast<const char*>(p);
    return {F::load(c), F::load(c + F::kBytes), F::load(c + 2 * F::kBytes),
            F::load(c + 3 * F::kBytes)};
  }
  __device__ __forceinline__ void store(void* p) const {
    char* c = reinterpret_cast<char*>(p);
    x.store(c);
    y.store(c + F::kBytes);
    zz.store(c + 2 * F::kBytes);
    zzz.store(c + 3 * F::kBytes);
  }
  __device__ __forceinline__ XYZZ neg() const { return {x, y.neg(), zz, zzz}; }

  // dbl-2008-s-1 (a = 0): 6M + 3S ... here S counted as M
  __device__ __forceinline__ XYZZ dbl() const {
    if (is_inf()) return *this;
    F u = y.dbl();
    F v = u.sqr();
    F w = u * v;
    F s = x * v;
    F xx = x.sqr();
    F m = xx.dbl() + xx;
    XYZZ r;
    r.x = m.sqr() - s.dbl();
    r.y = F::mul_sub_mul(m, s - r.x, w, y);
    r.zz = v * zz;
    r.zzz = w * zzz;
    return r;
  }

  // mixed addition  this += (px, py)   [madd-2008-s; doubling / cancellation handled]
  __device__ __forceinline__ void madd(const Affine<F>& p) {
    if (p.is_inf()) return;
    if (is_inf()) {
      x = p.x; y = p.y; zz = F::one(); zzz = F::one();
      return;
    }
    F pp = p.x * zz - x;     // P = U2 - X1
    F r = p.y * zzz - y;     // R = S2 - Y1
    if (pp.is_zero()) {
      if (r.is_zero()) {     // same point: mdbl-2008-s-1 on the affine operand
        F u = p.y.dbl();
        F v = u.sqr();
        F w = u * v;
        F s = p.x * v;
        F xx = p.x.sqr();
        F m = xx.dbl() + xx;
        x = m.sqr() - s.dbl();
        y = F::mul_sub_mul(m, s - x, w, p.y);
        zz = v;
        zzz = w;
      } else {
        *this = infinity();
      }
      return;
    }
    F p2 = pp.sqr();
    F p3 = pp * p2;
    F q = x * p2;
    zz = zz * p2;            // p2 dies here, before the four-operand fused product below
    zzz = zzz * p3;
    x = r.sqr() - p3 - q.dbl();
    y = F::mul_sub_mul(r, q - x, y, p3);
  }

  // full addition  this += o   [add-2008-s]
  __device__ __forceinline__ void add(const XYZZ& o) {
    if (o.is_inf()) return;
    if (is_inf()) {
      *this = o;
      return;
    }
    F u1 = x * o.zz;
    F u2 = o.x * zz;
    F s1 = y * o.zzz;
    F s2 = o.y * zzz;
    F pp = u2 - u1;
    F r = s2 - s1;
    if (pp.is_zero()) {
      if (r.is_zero()) {
        *this = dbl();
      } else {
        *this = infinity();
      }
      return;
    }
    F p2 = pp.sqr();
    F p3 = pp * p2;
    F q = u1 * p2;
    F x3 = r.sqr() - p3 - q.dbl();
    y = F::mul_sub_mul(r, q - x3, s1, p3);
    x = x3;
    zz = zz * o.zz * p2;
    zzz = zzz * o.zzz * p3;
  }

  // -> affine Montgomery (one inversion); infinity -> all-zero
  __device__ Affine<F> to_affine() const {
    if (is_inf()) return {F::zero(), F::zero()};
    F zi = zzz.inverse();          // 1/ZZZ
    F zi2 = (zi * zz).sqr();       // (ZZ/ZZZ)^2 = 1/ZZ   since ZZ^3 = ZZZ^2
    return {x * zi2, y * zi};
  }
};

// k * P for a small non-negative integer k (double-and-add, MSB first)
template <class F>
__device__ XYZZ<F> mul_small(const XYZZ<F>& p, uint32_t k) {
  XYZZ<F> r = XYZZ<F>::infinity();
  for (int i = 31 - __clz(k | 1); i >= 0; i--) {
    r = r.dbl();
    if ((k >> i) & 1) r.add(p);
  }
  if (k == 0) return XYZZ<F>::infinity();
  return r;
}

// 256-bit scalar (canonical little-endian words) times P
template <class F>
__device__ XYZZ<F> mul_scalar(const XYZZ<F>& p, const uint32_t (&k)[8]) {
  XYZZ<F> r = XYZZ<F>::infinity();
  for (int i = 255; i >= 0; i--) {
    r = r.dbl();
    if ((k[i >> 5] >> (i & 31)) & 1) r.add(p);
  }
  return r;
}

using G1Affine = Affine<Fq>;
using G2Affine = Affine<Fq2>;
using G1XYZZ = XYZZ<Fq>;
using G2XYZZ = XYZZ<Fq2>;
// storage sizes (identical for Fq2 and its lane-pair form Fq2L)
template <class F> constexpr size_t affine_bytes() { return 2 * size_t(F::kBytes); }
template <class F> constexpr size_t xyzz_bytes() { return 4 * size_t(F::kBytes); }

#endif
}  // namespace zkb
