// ff.cuh — BN254 Fq / Fr Montgomery arithmetic on 8x32-bit limbs for sm_100a.
//
// Replaces (there is no reference code for it — SURVEY.md §8(a) row a13): the field layer of
// the external prover eigen-zeth reaches through GenFinalProofRequest
// (reference proto/prover/v1/prover.proto:130-148).
//
// Representation: little-endian limbs, Montgomery form (x*2^256 mod p), *lazily reduced*:
// every value that crosses a function boundary is in [0, 2p).  Because p < 2^254, 4p < 2^256,
// so a+b never carries out of 256 bits and the Montgomery product of two values < 2p is again
// < 2p without a final subtraction:  (ab + mp)/2^256 < 4p^2/2^256 + p < 2p.
//
// The multiplier is a two-lane CIOS: partial products a_j*b_i with j even are 64-bit aligned on
// even limb positions, those with j odd on odd positions, so every `mad.lo.cc/madc.hi.cc` pair
// is one IMAD.WIDE.U32 with carry-in/out and each lane is a single uninterrupted carry chain.
// After the reduction step the low limb of the even lane is zero; dividing by 2^32 swaps the
// roles of the two lanes.  Every carry chain lives inside ONE asm statement, so the compiler
// can never schedule anything between a carry producer and its consumer.
#pragma once
#include <cstdint>

namespace zkb {

struct FqParams {
  static constexpr uint32_t M0 = 0xe4866389u;  // -q^{-1} mod 2^32
  __host__ __device__ static constexpr uint32_t mod(int i) {
    constexpr uint32_t m[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u,
                               0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
    return m[i];
  }
  __host__ __device__ static constexpr uint32_t mod2(int i) {
    constexpr uint32_t m[8] = {0xb0f9fa8eu, 0x7841182du, 0xd0e3951au, 0x2f02d522u,
                               0x0302b0bbu, 0x70a08b6du, 0xc2634053u, 0x60c89ce5u};
    return m[i];
  }
  __host__ __device__ static constexpr uint32_t one(int i) {  // 2^256 mod q
    constexpr uint32_t m[8] = {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u,
                               0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
    return m[i];
  }
  __host__ __device__ static constexpr uint32_t r2(int i) {  // 2^512 mod q
    constexpr uint32_t m[8] = {0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u,
                               0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u};
    return m[i];
  }
};

struct FrParams {
  static constexpr uint32_t M0 = 0xefffffffu;  // -r^{-1} mod 2^32
  __host__ __device__ static constexpr uint32_t mod(int i) {
    constexpr uint32_t m[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u,
                               0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
    return m[i];
  }
  __host__ __device__ static constexpr uint32_t mod2(int i) {
    constexpr uint32_t m[8] = {0xe0000002u, 0x87c3eb27u, 0xf372e122u, 0x5067d090u,
                               0x0302b0bau, 0x70a08b6du, 0xc2634053u, 0x60c89ce5u};
    return m[i];
  }
  __host__ __device__ static constexpr uint32_t one(int i) {  // 2^256 mod r
    constexpr uint32_t m[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u,
                               0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
    return m[i];
  }
  __host__ __device__ static constexpr uint32_t r2(int i) {  // 2^512 mod r
    constexpr uint32_t m[8] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u,
                               0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u};
    return m[i];
  }
};

#ifdef __CUDACC__

// ---------------------------------------------------------------- carry-chain building blocks
// lane[0..7] = {a0,a1,a2,a3} (x) b   as four disjoint 64-bit products (no carries needed)
__device__ __forceinline__ void lane_mul(uint32_t (&l)[8], uint32_t a0, uint32_t a1, uint32_t a2,
                                         uint32_t a3, uint32_t b) {
  asm("mul.lo.u32 %0, %8, %12;\n\t"
      "mul.hi.u32 %1, %8, %12;\n\t"
      "mul.lo.u32 %2, %9, %12;\n\t"
      "mul.hi.u32 %3, %9, %12;\n\t"
      "mul.lo.u32 %4, %10, %12;\n\t"
      "mul.hi.u32 %5, %10, %12;\n\t"
      "mul.lo.u32 %6, %11, %12;\n\t"
      "mul.hi.u32 %7, %11, %12;"
      : "=&r"(l[0]), "=&r"(l[1]), "=&r"(l[2]), "=&r"(l[3]), "=&r"(l[4]), "=&r"(l[5]), "=&r"(l[6]),
        "=&r"(l[7])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
}

// lane += {a0..a3} (x) b as one carry chain; the carry out of lane[7] is added to `spill`
// (spill is the top limb of the other lane, which sits one limb position higher).
__device__ __forceinline__ void lane_mad_spill(uint32_t (&l)[8], uint32_t& spill, uint32_t a0,
                                               uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b) {
  asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
      "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
      "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
      "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
      "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
      "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
      "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
      "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
      "addc.u32 %8, %8, 0;"
      : "+r"(l[0]), "+r"(l[1]), "+r"(l[2]), "+r"(l[3]), "+r"(l[4]), "+r"(l[5]), "+r"(l[6]),
        "+r"(l[7]), "+r"(spill)
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
}

// lane += {a0..a3} (x) b; the caller guarantees the sum fits (no carry out of lane[7]).
__device__ __forceinline__ void lane_mad(uint32_t (&l)[8], uint32_t a0, uint32_t a1, uint32_t a2,
                                         uint32_t a3, uint32_t b) {
  asm("mad.lo.cc.u32 %0, %8, %12, %0;\n\t"
      "madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
      "madc.lo.cc.u32 %2, %9, %12, %2;\n\t"
      "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
      "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
      "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
      "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
      "madc.hi.u32 %7, %11, %12, %7;"
      : "+r"(l[0]), "+r"(l[1]), "+r"(l[2]), "+r"(l[3]), "+r"(l[4]), "+r"(l[5]), "+r"(l[6]),
        "+r"(l[7])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
}

// The lane `y` has just been the even lane of a reduction step: y[0]==0 and, after the division
// by 2^32, y[1] sits at position 0 (it belongs to the new even lane `x`), y[2..7] at positions
// 1..6.  This fuses  x[0] += y[1]  (carry into position 1) with
//   y_new[0..7] = (y[2..7],0,0) + {a0..a3} (x) b      (positions 1..8, cannot overflow).
__device__ __forceinline__ void lane_shift_mad(uint32_t& x0, uint32_t (&y)[8], uint32_t a0,
                                               uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b) {
  asm("add.cc.u32 %0, %0, %2;\n\t"
      "madc.lo.cc.u32 %1, %9, %13, %3;\n\t"
      "madc.hi.cc.u32 %2, %9, %13, %4;\n\t"
      "madc.lo.cc.u32 %3, %10, %13, %5;\n\t"
      "madc.hi.cc.u32 %4, %10, %13, %6;\n\t"
      "madc.lo.cc.u32 %5, %11, %13, %7;\n\t"
      "madc.hi.cc.u32 %6, %11, %13, %8;\n\t"
      "madc.lo.cc.u32 %7, %12, %13, 0;\n\t"
      "madc.hi.u32 %8, %12, %13, 0;"
      : "+r"(x0), "+r"(y[0]), "+r"(y[1]), "+r"(y[2]), "+r"(y[3]), "+r"(y[4]), "+r"(y[5]),
        "+r"(y[6]), "+r"(y[7])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
}

// One Montgomery reduction step on (even lane e @ positions 0..7, odd lane o @ positions 1..8):
// adds m*p with m = e[0]*M0 so that e[0] becomes 0.
template <class P>
__device__ __forceinline__ void redc_step(uint32_t (&e)[8], uint32_t (&o)[8]) {
  uint32_t m = e[0] * P::M0;
  lane_mad(o, P::mod(1), P::mod(3), P::mod(5), P::mod(7), m);
  lane_mad_spill(e, o[7], P::mod(0), P::mod(2), P::mod(4), P::mod(6), m);
}

template <class P>
__device__ __forceinline__ void mont_mul(uint32_t (&r)[8], const uint32_t (&a)[8],
                                         const uint32_t (&b)[8]) {
  uint32_t e[8], o[8];
  lane_mul(e, a[0], a[2], a[4], a[6], b[0]);
  lane_mul(o, a[1], a[3], a[5], a[7], b[0]);
  redc_step<P>(e, o);
#pragma unroll
  for (int i = 1; i < 8; i += 2) {
    // odd i: o is the even-aligned lane, e the one being shifted
    lane_shift_mad(o[0], e, a[1], a[3], a[5], a[7], b[i]);
    lane_mad_spill(o, e[7], a[0], a[2], a[4], a[6], b[i]);
    redc_step<P>(o, e);
    if (i + 1 < 8) {
      lane_shift_mad(e[0], o, a[1], a[3], a[5], a[7], b[i + 1]);
      lane_mad_spill(e, o[7], a[0], a[2], a[4], a[6], b[i + 1]);
      redc_step<P>(e, o);
    }
  }
  // after i = 7 the lane with a zero low limb is o: result = e + (o >> 32)
  asm("add.cc.u32 %0, %8, %16;\n\t"
      "addc.cc.u32 %1, %9, %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32 %7, %15, 0;"
      : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]),
        "=&r"(r[7])
      : "r"(e[0]), "r"(e[1]), "r"(e[2]), "r"(e[3]), "r"(e[4]), "r"(e[5]), "r"(e[6]), "r"(e[7]),
        "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]));
}

// r = (a*b + c*d) / 2^256 mod p in ONE pass: both products are accumulated word by word into the
// same two lanes and share the eight reduction steps — 128 + 72 IMAD.WIDE instead of 2 x 136.
// With all operands <= 2p the running value stays below 5p + 1 < 2^256 (5p = 0.945 * 2^256), so no
// lane outgrows its 8-limb frame, and the result is < (8p^2 + 2^256 p) / 2^256 < 2.51 p: the caller
// folds it back under 2p.  Schedule and bounds are modelled word by word in tools/mul2_model.py.
template <class P>
__device__ __forceinline__ void mont_mul2(uint32_t (&r)[8], const uint32_t (&a)[8],
                                          const uint32_t (&b)[8], const uint32_t (&c)[8],
                                          const uint32_t (&d)[8]) {
  uint32_t e[8], o[8];
  lane_mul(e, a[0], a[2], a[4], a[6], b[0]);
  lane_mul(o, a[1], a[3], a[5], a[7], b[0]);
  lane_mad(o, c[1], c[3], c[5], c[7], d[0]);
  lane_mad_spill(e, o[7], c[0], c[2], c[4], c[6], d[0]);
  redc_step<P>(e, o);
#pragma unroll
  for (int i = 1; i < 8; i += 2) {
    lane_shift_mad(o[0], e, a[1], a[3], a[5], a[7], b[i]);
    lane_mad_spill(o, e[7], a[0], a[2], a[4], a[6], b[i]);
    lane_mad(e, c[1], c[3], c[5], c[7], d[i]);
    lane_mad_spill(o, e[7], c[0], c[2], c[4], c[6], d[i]);
    redc_step<P>(o, e);
    if (i + 1 < 8) {
      lane_shift_mad(e[0], o, a[1], a[3], a[5], a[7], b[i + 1]);
      lane_mad_spill(e, o[7], a[0], a[2], a[4], a[6], b[i + 1]);
      lane_mad(o, c[1], c[3], c[5], c[7], d[i + 1]);
      lane_mad_spill(e, o[7], c[0], c[2], c[4], c[6], d[i + 1]);
      redc_step<P>(e, o);
    }
  }
  asm("add.cc.u32 %0, %8, %16;\n\t"
      "addc.cc.u32 %1, %9, %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32 %7, %15, 0;"
      : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]),
        "=&r"(r[7])
      : "r"(e[0]), "r"(e[1]), "r"(e[2]), "r"(e[3]), "r"(e[4]), "r"(e[5]), "r"(e[6]), "r"(e[7]),
        "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]));
}

// r = a + b (256-bit, no carry out by the 4p < 2^256 invariant)
__device__ __forceinline__ void add256(uint32_t (&r)[8], const uint32_t (&a)[8],
                                       const uint32_t (&b)[8]) {
  asm("add.cc.u32 %0, %8, %16;\n\t"
      "addc.cc.u32 %1, %9, %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32 %7, %15, %23;"
      : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]),
        "=&r"(r[7])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
        "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
}

// r = a - b (256-bit), returns the borrow as 0 / 0xffffffff
__device__ __forceinline__ uint32_t sub256(uint32_t (&r)[8], const uint32_t (&a)[8],
                                           const uint32_t (&b)[8]) {
  uint32_t borrow;
  asm("sub.cc.u32 %0, %9, %17;\n\t"
      "subc.cc.u32 %1, %10, %18;\n\t"
      "subc.cc.u32 %2, %11, %19;\n\t"
      "subc.cc.u32 %3, %12, %20;\n\t"
      "subc.cc.u32 %4, %13, %21;\n\t"
      "subc.cc.u32 %5, %14, %22;\n\t"
      "subc.cc.u32 %6, %15, %23;\n\t"
      "subc.cc.u32 %7, %16, %24;\n\t"
      "subc.u32 %8, 0, 0;"
      : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]),
        "=&r"(r[7]), "=&r"(borrow)
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
        "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
  return borrow;
}

// ---------------------------------------------------------------- squaring
// a^2 as 16 words: the 28 off-diagonal products a_i a_j (i < j) are laid out in two lanes (odd and
// even i+j, so every chain walks 64-bit-aligned slots), summed, doubled with funnel shifts, and the
// 8 diagonal squares are added in one 16-word chain: 36 wide multiplications instead of 64.
// The schedule is modelled word by word in tools/sqr_model.py.
__device__ __forceinline__ void square_wide(uint32_t (&t)[16], const uint32_t (&a)[8]) {
  uint32_t o1, o2, o3, o4, o5, o6, o7, o8, o9, o10, o11, o12, o13, o14;
  uint32_t e2, e3, e4, e5, e6, e7, e8, e9, e10, e11, e12, e13;
  // odd lane, disjoint slots: (0,1)@1 (0,3)@3 (0,5)@5 (0,7)@7 (2,7)@9 (4,7)@11 (6,7)@13
  asm("mul.lo.u32 %0, %14, %15;\n\t mul.hi.u32 %1, %14, %15;\n\t"
      "mul.lo.u32 %2, %14, %16;\n\t mul.hi.u32 %3, %14, %16;\n\t"
      "mul.lo.u32 %4, %14, %17;\n\t mul.hi.u32 %5, %14, %17;\n\t"
      "mul.lo.u32 %6, %14, %18;\n\t mul.hi.u32 %7, %14, %18;\n\t"
      "mul.lo.u32 %8, %19, %18;\n\t mul.hi.u32 %9, %19, %18;\n\t"
      "mul.lo.u32 %10, %20, %18;\n\t mul.hi.u32 %11, %20, %18;\n\t"
      "mul.lo.u32 %12, %21, %18;\n\t mul.hi.u32 %13, %21, %18;"
      : "=&r"(o1), "=&r"(o2), "=&r"(o3), "=&r"(o4), "=&r"(o5), "=&r"(o6), "=&r"(o7), "=&r"(o8),
        "=&r"(o9), "=&r"(o10), "=&r"(o11), "=&r"(o12), "=&r"(o13), "=&r"(o14)
      : "r"(a[0]), "r"(a[1]), "r"(a[3]), "r"(a[5]), "r"(a[7]), "r"(a[2]), "r"(a[4]), "r"(a[6]));
  // (1,2)@3 (1,4)@5 (1,6)@7 (3,6)@9 (5,6)@11, carry rippled to o14
  asm("mad.lo.cc.u32 %0, %12, %13, %0;\n\t madc.hi.cc.u32 %1, %12, %13, %1;\n\t"
      "madc.lo.cc.u32 %2, %12, %14, %2;\n\t madc.hi.cc.u32 %3, %12, %14, %3;\n\t"
      "madc.lo.cc.u32 %4, %12, %15, %4;\n\t madc.hi.cc.u32 %5, %12, %15, %5;\n\t"
      "madc.lo.cc.u32 %6, %16, %15, %6;\n\t madc.hi.cc.u32 %7, %16, %15, %7;\n\t"
      "madc.lo.cc.u32 %8, %17, %15, %8;\n\t madc.hi.cc.u32 %9, %17, %15, %9;\n\t"
      "addc.cc.u32 %10, %10, 0;\n\t addc.u32 %11, %11, 0;"
      : "+r"(o3), "+r"(o4), "+r"(o5), "+r"(o6), "+r"(o7), "+r"(o8), "+r"(o9), "+r"(o10), "+r"(o11),
        "+r"(o12), "+r"(o13), "+r"(o14)
      : "r"(a[1]), "r"(a[2]), "r"(a[4]), "r"(a[6]), "r"(a[3]), "r"(a[5]));
  // (2,3)@5 (2,5)@7 (4,5)@9
  asm("mad.lo.cc.u32 %0, %10, %11, %0;\n\t madc.hi.cc.u32 %1, %10, %11, %1;\n\t"
      "madc.lo.cc.u32 %2, %10, %12, %2;\n\t madc.hi.cc.u32 %3, %10, %12, %3;\n\t"
      "madc.lo.cc.u32 %4, %13, %12, %4;\n\t madc.hi.cc.u32 %5, %13, %12, %5;\n\t"
      "addc.cc.u32 %6, %6, 0;\n\t addc.cc.u32 %7, %7, 0;\n\t addc.cc.u32 %8, %8, 0;\n\t"
      "addc.u32 %9, %9, 0;"
      : "+r"(o5), "+r"(o6), "+r"(o7), "+r"(o8), "+r"(o9), "+r"(o10), "+r"(o11), "+r"(o12), "+r"(o13),
        "+r"(o14)
      : "r"(a[2]), "r"(a[3]), "r"(a[5]), "r"(a[4]));
  // (3,4)@7
  asm("mad.lo.cc.u32 %0, %8, %9, %0;\n\t madc.hi.cc.u32 %1, %8, %9, %1;\n\t"
      "addc.cc.u32 %2, %2, 0;\n\t addc.cc.u32 %3, %3, 0;\n\t addc.cc.u32 %4, %4, 0;\n\t"
      "addc.cc.u32 %5, %5, 0;\n\t addc.cc.u32 %6, %6, 0;\n\t addc.u32 %7, %7, 0;"
      : "+r"(o7), "+r"(o8), "+r"(o9), "+r"(o10), "+r"(o11), "+r"(o12), "+r"(o13), "+r"(o14)
      : "r"(a[3]), "r"(a[4]));
  // even lane, disjoint slots: (0,2)@2 (0,4)@4 (0,6)@6 (1,7)@8 (3,7)@10 (5,7)@12
  asm("mul.lo.u32 %0, %12, %13;\n\t mul.hi.u32 %1, %12, %13;\n\t"
      "mul.lo.u32 %2, %12, %14;\n\t mul.hi.u32 %3, %12, %14;\n\t"
      "mul.lo.u32 %4, %12, %15;\n\t mul.hi.u32 %5, %12, %15;\n\t"
      "mul.lo.u32 %6, %16, %17;\n\t mul.hi.u32 %7, %16, %17;\n\t"
      "mul.lo.u32 %8, %18, %17;\n\t mul.hi.u32 %9, %18, %17;\n\t"
      "mul.lo.u32 %10, %19, %17;\n\t mul.hi.u32 %11, %19, %17;"
      : "=&r"(e2), "=&r"(e3), "=&r"(e4), "=&r"(e5), "=&r"(e6), "=&r"(e7), "=&r"(e8), "=&r"(e9),
        "=&r"(e10), "=&r"(e11), "=&r"(e12), "=&r"(e13)
      : "r"(a[0]), "r"(a[2]), "r"(a[4]), "r"(a[6]), "r"(a[1]), "r"(a[7]), "r"(a[3]), "r"(a[5]));
  // (1,3)@4 (1,5)@6 (2,6)@8 (4,6)@10
  asm("mad.lo.cc.u32 %0, %10, %11, %0;\n\t madc.hi.cc.u32 %1, %10, %11, %1;\n\t"
      "madc.lo.cc.u32 %2, %10, %12, %2;\n\t madc.hi.cc.u32 %3, %10, %12, %3;\n\t"
      "madc.lo.cc.u32 %4, %13, %14, %4;\n\t madc.hi.cc.u32 %5, %13, %14, %5;\n\t"
      "madc.lo.cc.u32 %6, %15, %14, %6;\n\t madc.hi.cc.u32 %7, %15, %14, %7;\n\t"
      "addc.cc.u32 %8, %8, 0;\n\t addc.u32 %9, %9, 0;"
      : "+r"(e4), "+r"(e5), "+r"(e6), "+r"(e7), "+r"(e8), "+r"(e9), "+r"(e10), "+r"(e11), "+r"(e12),
        "+r"(e13)
      : "r"(a[1]), "r"(a[3]), "r"(a[5]), "r"(a[2]), "r"(a[6]), "r"(a[4]));
  // (2,4)@6 (3,5)@8
  asm("mad.lo.cc.u32 %0, %8, %9, %0;\n\t madc.hi.cc.u32 %1, %8, %9, %1;\n\t"
      "madc.lo.cc.u32 %2, %10, %11, %2;\n\t madc.hi.cc.u32 %3, %10, %11, %3;\n\t"
      "addc.cc.u32 %4, %4, 0;\n\t addc.cc.u32 %5, %5, 0;\n\t addc.cc.u32 %6, %6, 0;\n\t"
      "addc.u32 %7, %7, 0;"
      : "+r"(e6), "+r"(e7), "+r"(e8), "+r"(e9), "+r"(e10), "+r"(e11), "+r"(e12), "+r"(e13)
      : "r"(a[2]), "r"(a[4]), "r"(a[3]), "r"(a[5]));
  // S = O + E (into the odd lane, words 2..14)
  asm("add.cc.u32 %0, %0, %13;\n\t addc.cc.u32 %1, %1, %14;\n\t addc.cc.u32 %2, %2, %15;\n\t"
      "addc.cc.u32 %3, %3, %16;\n\t addc.cc.u32 %4, %4, %17;\n\t addc.cc.u32 %5, %5, %18;\n\t"
      "addc.cc.u32 %6, %6, %19;\n\t addc.cc.u32 %7, %7, %20;\n\t addc.cc.u32 %8, %8, %21;\n\t"
      "addc.cc.u32 %9, %9, %22;\n\t addc.cc.u32 %10, %10, %23;\n\t addc.cc.u32 %11, %11, %24;\n\t"
      "addc.u32 %12, %12, 0;"
      : "+r"(o2), "+r"(o3), "+r"(o4), "+r"(o5), "+r"(o6), "+r"(o7), "+r"(o8), "+r"(o9), "+r"(o10),
        "+r"(o11), "+r"(o12), "+r"(o13), "+r"(o14)
      : "r"(e2), "r"(e3), "r"(e4), "r"(e5), "r"(e6), "r"(e7), "r"(e8), "r"(e9), "r"(e10), "r"(e11),
        "r"(e12), "r"(e13));
  // T = 2 S ...
  t[1] = o1 << 1;
  t[2] = __funnelshift_l(o1, o2, 1);
  t[3] = __funnelshift_l(o2, o3, 1);
  t[4] = __funnelshift_l(o3, o4, 1);
  t[5] = __funnelshift_l(o4, o5, 1);
  t[6] = __funnelshift_l(o5, o6, 1);
  t[7] = __funnelshift_l(o6, o7, 1);
  t[8] = __funnelshift_l(o7, o8, 1);
  t[9] = __funnelshift_l(o8, o9, 1);
  t[10] = __funnelshift_l(o9, o10, 1);
  t[11] = __funnelshift_l(o10, o11, 1);
  t[12] = __funnelshift_l(o11, o12, 1);
  t[13] = __funnelshift_l(o12, o13, 1);
  t[14] = __funnelshift_l(o13, o14, 1);
  t[15] = o14 >> 31;
  // ... + the diagonal a_i^2 at word 2i, one chain over all 16 words
  t[0] = a[0] * a[0];
  asm("mad.hi.cc.u32 %0, %15, %15, %0;\n\t"
      "madc.lo.cc.u32 %1, %16, %16, %1;\n\t madc.hi.cc.u32 %2, %16, %16, %2;\n\t"
      "madc.lo.cc.u32 %3, %17, %17, %3;\n\t madc.hi.cc.u32 %4, %17, %17, %4;\n\t"
      "madc.lo.cc.u32 %5, %18, %18, %5;\n\t madc.hi.cc.u32 %6, %18, %18, %6;\n\t"
      "madc.lo.cc.u32 %7, %19, %19, %7;\n\t madc.hi.cc.u32 %8, %19, %19, %8;\n\t"
      "madc.lo.cc.u32 %9, %20, %20, %9;\n\t madc.hi.cc.u32 %10, %20, %20, %10;\n\t"
      "madc.lo.cc.u32 %11, %21, %21, %11;\n\t madc.hi.cc.u32 %12, %21, %21, %12;\n\t"
      "madc.lo.cc.u32 %13, %22, %22, %13;\n\t madc.hi.u32 %14, %22, %22, %14;"
      : "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7]), "+r"(t[8]),
        "+r"(t[9]), "+r"(t[10]), "+r"(t[11]), "+r"(t[12]), "+r"(t[13]), "+r"(t[14]), "+r"(t[15])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]));
}

// a^2 / 2^256 mod p, lazily reduced:  T = a^2 = T_hi 2^256 + T_lo;  U = (T_lo + m p) / 2^256 <= p by
// eight reduction rounds on the two lanes (the mont_mul schedule without the a*b_i terms);
// result = T_hi + U < 4p^2/2^256 + p + 1 <= 2p.
template <class P>
__device__ __forceinline__ void mont_sqr(uint32_t (&r)[8], const uint32_t (&a)[8]) {
  uint32_t t[16];
  square_wide(t, a);
  uint32_t e[8], o[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { e[i] = t[i]; o[i] = 0; }
  redc_step<P>(e, o);
#pragma unroll
  for (int i = 1; i < 8; i += 2) {
    {
      uint32_t m = (o[0] + e[1]) * P::M0;
      lane_shift_mad(o[0], e, P::mod(1), P::mod(3), P::mod(5), P::mod(7), m);
      lane_mad_spill(o, e[7], P::mod(0), P::mod(2), P::mod(4), P::mod(6), m);
    }
    if (i + 1 < 8) {
      uint32_t m = (e[0] + o[1]) * P::M0;
      lane_shift_mad(e[0], o, P::mod(1), P::mod(3), P::mod(5), P::mod(7), m);
      lane_mad_spill(e, o[7], P::mod(0), P::mod(2), P::mod(4), P::mod(6), m);
    }
  }
  // U = e + (o >> 32), then r = U + T_hi
  uint32_t u[8];
  asm("add.cc.u32 %0, %8, %16;\n\t"
      "addc.cc.u32 %1, %9, %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32 %7, %15, 0;"
      : "=&r"(u[0]), "=&r"(u[1]), "=&r"(u[2]), "=&r"(u[3]), "=&r"(u[4]), "=&r"(u[5]), "=&r"(u[6]),
        "=&r"(u[7])
      : "r"(e[0]), "r"(e[1]), "r"(e[2]), "r"(e[3]), "r"(e[4]), "r"(e[5]), "r"(e[6]), "r"(e[7]),
        "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]));
  uint32_t hi[8];
#pragma unroll
  for (int i = 0; i < 8; i++) hi[i] = t[8 + i];
  add256(r, u, hi);
}

// ---------------------------------------------------------------- the field type
template <class P>
struct Fp {
  uint32_t v[8];
  using Params = P;
  static constexpr int kBytes = 32;   // storage size
  static constexpr int kLanes = 1;    // threads that hold one element

  __device__ __forceinline__ static Fp zero() {
    Fp r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = 0;
    return r;
  }
  __device__ __forceinline__ static Fp one() {
    Fp r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = P::one(i);
    return r;
  }
  __device__ __forceinline__ static Fp modulus() {
    Fp r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = P::mod(i);
    return r;
  }
  __device__ __forceinline__ static Fp modulus2() {
    Fp r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = P::mod2(i);
    return r;
  }
  __device__ __forceinline__ static Fp r2() {
    Fp r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = P::r2(i);
    return r;
  }

  // 128-bit vectorised global/shared access (pointer must be 16-byte aligned)
  __device__ __forceinline__ static Fp load(const void* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 lo = q[0], hi = q[1];
    Fp r;
    r.v[0] = lo.x; r.v[1] = lo.y; r.v[2] = lo.z; r.v[3] = lo.w;
    r.v[4] = hi.x; r.v[5] = hi.y; r.v[6] = hi.z; r.v[7] = hi.w;
    return r;
  }
  __device__ __forceinline__ static Fp load_nc(const void* p) {  // read-only path
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 lo = __ldg(q), hi = __ldg(q + 1);
    Fp r;
    r.v[0] = lo.x; r.v[1] = lo.y; r.v[2] = lo.z; r.v[3] = lo.w;
    r.v[4] = hi.x; r.v[5] = hi.y; r.v[6] = hi.z; r.v[7] = hi.w;
    return r;
  }
  __device__ __forceinline__ void store(void* p) const {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(v[0], v[1], v[2], v[3]);
    q[1] = make_uint4(v[4], v[5], v[6], v[7]);
  }

  // [0,4p) -> [0,2p)
  __device__ __forceinline__ void fold2p() {
    uint32_t t[8];
    uint32_t m2[8];
#pragma unroll
    for (int i = 0; i < 8; i++) m2[i] = P::mod2(i);
    uint32_t borrow = sub256(t, v, m2);
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = borrow ? v[i] : t[i];
  }
  // [0,2p) -> [0,p)   (canonical residue)
  __device__ __forceinline__ Fp reduced() const {
    uint32_t t[8];
    uint32_t m[8];
#pragma unroll
    for (int i = 0; i < 8; i++) m[i] = P::mod(i);
    uint32_t borrow = sub256(t, v, m);
    Fp r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = borrow ? v[i] : t[i];
    return r;
  }

  __device__ __forceinline__ friend Fp operator+(const Fp& a, const Fp& b) {
    Fp r;
    add256(r.v, a.v, b.v);
    r.fold2p();
    return r;
  }
  __device__ __forceinline__ friend Fp operator-(const Fp& a, const Fp& b) {
    Fp r;
    uint32_t borrow = sub256(r.v, a.v, b.v);
    uint32_t m2[8];
#pragma unroll
    for (int i = 0; i < 8; i++) m2[i] = P::mod2(i) & borrow;
    uint32_t t[8];
    add256(t, r.v, m2);
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = t[i];
    return r;
  }
  __device__ __forceinline__ friend Fp operator*(const Fp& a, const Fp& b) {
    Fp r;
    mont_mul<P>(r.v, a.v, b.v);
    return r;
  }
  // a*b + c*d and a*b - c*d with one shared Montgomery reduction (mont_mul2), result in [0, 2p)
  __device__ __forceinline__ static Fp mul_add_mul(const Fp& a, const Fp& b, const Fp& c, const Fp& d) {
    Fp r;
    mont_mul2<P>(r.v, a.v, b.v, c.v, d.v);
    r.fold2p();
    return r;
  }
  __device__ __forceinline__ static Fp mul_sub_mul(const Fp& a, const Fp& b, const Fp& c, const Fp& d) {
    return mul_add_mul(a, b, zero() - c, d);
  }
  // The hot kernels square with the general multiplier: the dedicated squaring below needs 19 %
  // fewer IMAD.WIDE but its carry ripples, funnel shifts and extra live registers cancel the gain
  // at the 128-register cap of the MSM walk (measured on B200: 33.65 ms vs 33.1 ms at 2^24).
  __device__ __forceinline__ Fp sqr() const { return (*this) * (*this); }
  __device__ __forceinline__ Fp sqr_dedicated() const {
    Fp r;
    mont_sqr<P>(r.v, v);
    return r;
  }
  __device__ __forceinline__ Fp dbl() const { return (*this) + (*this); }
  __device__ __forceinline__ Fp neg() const { return zero() - (*this); }

  // value == 0 mod p for a lazily reduced value: the only representatives are 0 and p
  __device__ __forceinline__ bool is_zero() const {
    uint32_t o = 0, x = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      o |= v[i];
      x |= v[i] ^ P::mod(i);
    }
    return o == 0 || x == 0;
  }
  __device__ __forceinline__ bool raw_is_zero() const {  // all limbs zero (encoding tests)
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= v[i];
    return o == 0;
  }
  __device__ __forceinline__ friend bool operator==(const Fp& a, const Fp& b) {
    return (a - b).is_zero();
  }

  // canonical integer <-> Montgomery
  __device__ __forceinline__ Fp to_mont() const { return (*this) * r2(); }
  __device__ __forceinline__ Fp from_mont() const {
    Fp o = zero();
    o.v[0] = 1;
    return ((*this) * o).reduced();
  }

  // a^e for a 256-bit little-endian exponent (not constant time; used for inversion only)
  __device__ Fp pow(const uint32_t (&e)[8]) const {
    Fp r = one();
    for (int i = 255; i >= 0; i--) {
      r = r.sqr();
      if ((e[i >> 5] >> (i & 31)) & 1) r = r * (*this);
    }
    return r;
  }
  // Fermat inversion; 0 -> 0 (kept as the cross-check of inverse(): zkb_selftest field op)
  __device__ Fp inverse_fermat() const {
    uint32_t e[8];
#pragma unroll
    for (int i = 0; i < 8; i++) e[i] = P::mod(i);
    e[0] -= 2;  // both moduli have low limb >= 2
    return pow(e);
  }
  // Inversion by the binary extended Euclid on the raw residue (shifts, additions and subtractions of
  // 256-bit integers only): ~360 loop turns of ~60 instructions, against 256 squarings + ~110
  // products of ~270 instructions each for Fermat.  An inversion is a single-thread dependent chain
  // at the end of every MSM tail; this form takes a third of the time.  Not constant time (neither
  // was pow()); 0 -> 0.  With x = aR the loop yields (aR)^-1, and one product by R^3 brings it back to
  // a^-1 R.
  __device__ Fp inverse() const {
    const Fp a = reduced();
    if (a.raw_is_zero()) return zero();
    uint32_t u[8], w[8], x1[8], x2[8], m[8], t[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
      u[i] = a.v[i];
      w[i] = m[i] = P::mod(i);
      x1[i] = i == 0;
      x2[i] = 0;
    }
    auto is_one = [](const uint32_t (&x)[8]) {
      uint32_t o = x[0] ^ 1u;
#pragma unroll
      for (int i = 1; i < 8; i++) o |= x[i];
      return o == 0;
    };
    auto shr1 = [](uint32_t (&x)[8]) {
#pragma unroll
      for (int i = 0; i < 7; i++) x[i] = __funnelshift_r(x[i], x[i + 1], 1);
      x[7] >>= 1;
    };
    auto halve_mod = [&](uint32_t (&x)[8]) {          // x / 2 mod p, x < p (x + p < 2^255)
      if (x[0] & 1u) {
        add256(t, x, m);
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = t[i];
      }
      shr1(x);
    };
    auto sub_mod = [&](uint32_t (&x)[8], const uint32_t (&y)[8]) {   // x = x - y mod p
      uint32_t t2[8];
      const uint32_t borrow = sub256(t, x, y);
      if (borrow) {
        add256(t2, t, m);
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = t2[i];
      } else {
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = t[i];
      }
    };
    while (!is_one(u) && !is_one(w)) {
      while (!(u[0] & 1u)) {
        shr1(u);
        halve_mod(x1);
      }
      while (!(w[0] & 1u)) {
        shr1(w);
        halve_mod(x2);
      }
      uint32_t d[8];
      const uint32_t borrow = sub256(d, u, w);
      if (!borrow) {                                  // u >= w
#pragma unroll
        for (int i = 0; i < 8; i++) u[i] = d[i];
        sub_mod(x1, x2);
      } else {
        sub256(d, w, u);
#pragma unroll
        for (int i = 0; i < 8; i++) w[i] = d[i];
        sub_mod(x2, x1);
      }
    }
    Fp r;
    const bool from_u = is_one(u);
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = from_u ? x1[i] : x2[i];
    return r * (r2() * r2());
  }
};

using Fq = Fp<FqParams>;
using Fr = Fp<FrParams>;

// ---------------------------------------------------------------- Fq2 = Fq[u]/(u^2+1)
// CALLS = false: every product is inlined (the bucket walk: straight-line code, best pipe use).
// CALLS = true: products and squares are ONE out-of-line routine each (operands and result in
// registers, no stack).  The latency-bound G2 kernels around the walk (partial merges, bucket
// reduction, window sums, Horner) inline 15-60 products per loop turn otherwise — hundreds of KB of
// straight-line code whose instruction-cache misses showed up as `stalled_no_instruction` in ncu;
// with the call form the G2 bucket-reduction tail went from 2.16 to 1.73 ms (2^19 points), while the
// walk itself is 3 % FASTER inlined (the call moves cost more than the cache misses), so it stays so.
template <bool CALLS>
struct Fq2T {
  Fq c0, c1;
  static constexpr int kBytes = 64;
  static constexpr int kLanes = 1;
  __device__ __forceinline__ static Fq2T zero() { return {Fq::zero(), Fq::zero()}; }
  __device__ __forceinline__ static Fq2T one() { return {Fq::one(), Fq::zero()}; }
  __device__ __forceinline__ static Fq2T load(const void* p) {
    const char* c = reinterpret_cast<const char*>(p);
    return {Fq::load(c), Fq::load(c + 32)};
  }
  __device__ __forceinline__ static Fq2T load_nc(const void* p) {
    const char* c = reinterpret_cast<const char*>(p);
    return {Fq::load_nc(c), Fq::load_nc(c + 32)};
  }
  __device__ __forceinline__ void store(void* p) const {
    char* c = reinterpret_cast<char*>(p);
    c0.store(c);
    c1.store(c + 32);
  }
  __device__ __forceinline__ friend Fq2T operator+(const Fq2T& a, const Fq2T& b) {
    return {a.c0 + b.c0, a.c1 + b.c1};
  }
  __device__ __forceinline__ friend Fq2T operator-(const Fq2T& a, const Fq2T& b) {
    return {a.c0 - b.c0, a.c1 - b.c1};
  }
  // Two fused dual products (a0 b0 - a1 b1, a0 b1 + a1 b0): 2 x 200 IMAD.WIDE, against 3 x 136 for
  // Karatsuba — and without Karatsuba's five 256-bit additions / subtractions and their temporaries.
  __device__ __forceinline__ static Fq2T mul_inline(const Fq2T& a, const Fq2T& b) {
    return {Fq::mul_sub_mul(a.c0, b.c0, a.c1, b.c1), Fq::mul_add_mul(a.c0, b.c1, a.c1, b.c0)};
  }
  // (a0+a1)(a0-a1), 2 a0 a1 : 2 Fq multiplications
  __device__ __forceinline__ static Fq2T sqr_inline(const Fq2T& a) {
    Fq t = a.c0 * a.c1;
    return {(a.c0 + a.c1) * (a.c0 - a.c1), t.dbl()};
  }
  __device__ __noinline__ static Fq2T mul_call(Fq2T a, Fq2T b) { return mul_inline(a, b); }
  __device__ __noinline__ static Fq2T sqr_call(Fq2T a) { return sqr_inline(a); }
  __device__ __forceinline__ friend Fq2T operator*(const Fq2T& a, const Fq2T& b) {
    if constexpr (CALLS) return mul_call(a, b);
    else return mul_inline(a, b);
  }
  __device__ __forceinline__ static Fq2T mul_sub_mul(const Fq2T& a, const Fq2T& b, const Fq2T& c, const Fq2T& d) {
    return a * b - c * d;
  }
  __device__ __forceinline__ Fq2T sqr() const {
    if constexpr (CALLS) return sqr_call(*this);
    else return sqr_inline(*this);
  }
  __device__ __forceinline__ Fq2T dbl() const { return {c0.dbl(), c1.dbl()}; }
  __device__ __forceinline__ Fq2T neg() const { return {c0.neg(), c1.neg()}; }
  __device__ __forceinline__ bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
  __device__ __forceinline__ bool raw_is_zero() const {
    return c0.raw_is_zero() && c1.raw_is_zero();
  }
  __device__ __forceinline__ friend bool operator==(const Fq2T& a, const Fq2T& b) {
    return (a - b).is_zero();
  }
  __device__ __forceinline__ Fq2T to_mont() const { return {c0.to_mont(), c1.to_mont()}; }
  __device__ __forceinline__ Fq2T from_mont() const { return {c0.from_mont(), c1.from_mont()}; }
  __device__ __forceinline__ Fq2T reduced() const { return {c0.reduced(), c1.reduced()}; }
  __device__ Fq2T inverse() const {
    Fq d = (c0.sqr() + c1.sqr()).inverse();
    return {c0 * d, (c1 * d).neg()};
  }
};
using Fq2 = Fq2T<false>;
using Fq2C = Fq2T<true>;

// ---------------------------------------------------------------- Fq2 split over a lane pair
// One Fq2 element held by TWO adjacent lanes of a warp: the even lane owns c0, the odd lane c1.
// The memory layout is that of Fq2 (c0 | c1, 64 bytes); only the register distribution differs.
// Why: an XYZZ accumulator over Fq2 plus a prefetched point needs > 200 registers in one thread, which
// caps the G2 bucket walk at 8 warps per SM and leaves the IMAD.WIDE pipe idle a third of the time
// (2 warps per scheduler cannot cover a 136-deep carry chain).  Split over a pair, each lane carries
// exactly what a G1 lane carries (one Fq per coordinate) and the walk runs at 16 warps per SM.
//   (a0 + a1 u)(b0 + b1 u):  even lane  a0 b0 + (-a1) b1      odd lane  a1 b0 + a0 b1
// both are ONE fused dual product (mont_mul2) on  own_a * P + recv * Q  with recv = the partner's
// a-component (the odd lane sends -a1) and (P, Q) = (b, b') / (b', b): 16 shuffles and a few selects
// per Fq2 product, against 200 IMAD.WIDE.  Squaring: even (a0 + a1)(a0 - a1), odd (a1 + a1) a0.
// All shuffles name only the two lanes of the pair, which always take the same branches, so pairs of
// one warp may diverge from each other freely.  Block sizes must be even (they are multiples of 32).
__device__ __forceinline__ uint32_t pair_half() { return threadIdx.x & 1u; }
__device__ __forceinline__ uint32_t pair_mask() { return 3u << (threadIdx.x & 30u); }
__device__ __forceinline__ Fq pair_swap(const Fq& a) {
  Fq r;
  const uint32_t m = pair_mask();
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = __shfl_xor_sync(m, a.v[i], 1);
  return r;
}
__device__ __forceinline__ bool pair_all(bool x) {
  // the shuffle must be executed by BOTH lanes whatever x is (no short-circuit in front of it)
  const int other = __shfl_xor_sync(pair_mask(), int(x), 1);
  return x & (other != 0);
}
__device__ __forceinline__ Fq fq_select(bool c, const Fq& a, const Fq& b) {   // c ? a : b, branch-free
  Fq r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = c ? a.v[i] : b.v[i];
  return r;
}
// 2p - a  (a in [0, 2p]  ->  result in [0, 2p]: a valid multiplier operand of mont_mul / mont_mul2)
__device__ __forceinline__ Fq fq_neg_2p(const Fq& a) {
  Fq r;
  sub256(r.v, Fq::modulus2().v, a.v);
  return r;
}

struct Fq2L {
  Fq v;                               // c0 on even lanes, c1 on odd lanes
  static constexpr int kBytes = 64;
  static constexpr int kLanes = 2;
  __device__ __forceinline__ static Fq2L zero() { return {Fq::zero()}; }
  __device__ __forceinline__ static Fq2L one() { return {fq_select(pair_half() != 0, Fq::zero(), Fq::one())}; }
  __device__ __forceinline__ static Fq2L load(const void* p) {
    return {Fq::load(reinterpret_cast<const char*>(p) + 32 * pair_half())};
  }
  __device__ __forceinline__ static Fq2L load_nc(const void* p) {
    return {Fq::load_nc(reinterpret_cast<const char*>(p) + 32 * pair_half())};
  }
  __device__ __forceinline__ void store(void* p) const { v.store(reinterpret_cast<char*>(p) + 32 * pair_half()); }
  __device__ __forceinline__ friend Fq2L operator+(const Fq2L& a, const Fq2L& b) { return {a.v + b.v}; }
  __device__ __forceinline__ friend Fq2L operator-(const Fq2L& a, const Fq2L& b) { return {a.v - b.v}; }
  __device__ __forceinline__ friend Fq2L operator*(const Fq2L& a, const Fq2L& b) {
    const bool odd = pair_half() != 0;
    const Fq send = fq_select(odd, fq_neg_2p(a.v), a.v);     // the odd lane hands over -a1
    const Fq recv = pair_swap(send);                          // even: -a1      odd: a0
    const Fq bp = pair_swap(b.v);                             // the partner's b component
    const Fq P = fq_select(odd, bp, b.v);                     // even: b0       odd: b0
    const Fq Q = fq_select(odd, b.v, bp);                     // even: b1       odd: b1
    return {Fq::mul_add_mul(a.v, P, recv, Q)};                // even: a0 b0 - a1 b1   odd: a1 b0 + a0 b1
  }
  __device__ __forceinline__ static Fq2L mul_sub_mul(const Fq2L& a, const Fq2L& b, const Fq2L& c, const Fq2L& d) {
    return a * b - c * d;
  }
  __device__ __forceinline__ Fq2L sqr() const {
    const bool odd = pair_half() != 0;
    const Fq o = pair_swap(v);
    const Fq U = v + fq_select(odd, v, o);                    // even: a0 + a1   odd: 2 a1
    const Fq V = fq_select(odd, o, v - o);                    // even: a0 - a1   odd: a0
    return {U * V};
  }
  __device__ __forceinline__ Fq2L dbl() const { return {v.dbl()}; }
  __device__ __forceinline__ Fq2L neg() const { return {v.neg()}; }
  __device__ __forceinline__ bool is_zero() const { return pair_all(v.is_zero()); }
  __device__ __forceinline__ bool raw_is_zero() const { return pair_all(v.raw_is_zero()); }
  __device__ __forceinline__ friend bool operator==(const Fq2L& a, const Fq2L& b) { return (a - b).is_zero(); }
  __device__ __forceinline__ Fq2L to_mont() const { return {v.to_mont()}; }
  __device__ __forceinline__ Fq2L from_mont() const { return {v.from_mont()}; }
  __device__ __forceinline__ Fq2L reduced() const { return {v.reduced()}; }
  __device__ Fq2L inverse() const {
    const Fq sq = v.sqr();
    const Fq d = (sq + pair_swap(sq)).inverse();              // 1 / (c0^2 + c1^2), computed by both lanes
    const Fq r = v * d;
    return {fq_select(pair_half() != 0, r.neg(), r)};
  }
};

#endif  // __CUDACC__
}  // namespace zkb
