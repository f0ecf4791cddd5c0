// ntt.cu — Fr radix-2 NTT / iNTT / coset variants for BN254 on sm_100a.
//
// No reference counterpart exists.  Semantics follow ark-poly's
// Radix2EvaluationDomain (natural order in, natural order out, inverse carries 1/n, coset shift
// g = 5), the algorithm the north_star names as the CPU baseline.
//
// Decomposition: log2 n = k_0 + ... + k_{P-1}, k_p <= 8.  Pass p performs all size-2^{k_p}
// sub-transforms ("four-step" recursion, decimation in frequency):
//   * passes 0..P-2 ("column" passes) work in place on elements  blk*M + t*S + r  (t = transform
//     index, S = M >> k_p) and multiply the result K by the inter-pass twiddle w_M^{r*K};
//   * the last pass transforms contiguous rows and scatters them to natural order (digit
//     reversal of the row number), C adjacent outputs at a time.
// A CTA owns a tile of 2048 elements (64 KB of shared memory): 8 warps x 256 slots.  Each warp
// runs 2^{8-k} independent size-2^k transforms on its 256 slots, 8 elements per thread in
// registers, three register stage-groups with two shared-memory re-distributions in between.
// Shared memory is split in two 16-byte planes with an XOR swizzle so that every LDS.128/STS.128
// quarter-warp hits 8 distinct 16-byte bank groups.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <mutex>

#include "ff.cuh"
#include "ntt.h"

namespace zkb {

// ---------------------------------------------------------------- shared-memory geometry
constexpr int kWarpsPerCta = 8;
constexpr int kThreads = kWarpsPerCta * 32;
constexpr int kSlotsPerWarp = 256;
constexpr int kTileElems = kWarpsPerCta * kSlotsPerWarp;   // 2048
constexpr int kWarpPlaneStride = kSlotsPerWarp + 1;        // in uint4 units (+1: bank shift per warp)
constexpr int kPlaneStride = kWarpsPerCta * kWarpPlaneStride + 4;
constexpr int kDataU4 = 2 * kPlaneStride;
constexpr int kStageTw = 256;                              // 255 used
constexpr int kSmemU4 = kDataU4 + 2 * kStageTw;
constexpr size_t kSmemBytes = size_t(kSmemU4) * 16;

__device__ __forceinline__ int swz(int u) { return u ^ ((u >> 3) & 7); }

__device__ __forceinline__ void sm_store(uint4* sm, int warp, int u, const Fr& x) {
  int o = warp * kWarpPlaneStride + swz(u);
  sm[o] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
  sm[kPlaneStride + o] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
}
__device__ __forceinline__ Fr sm_load(const uint4* sm, int warp, int u) {
  int o = warp * kWarpPlaneStride + swz(u);
  uint4 lo = sm[o], hi = sm[kPlaneStride + o];
  Fr x;
  x.v[0] = lo.x; x.v[1] = lo.y; x.v[2] = lo.z; x.v[3] = lo.w;
  x.v[4] = hi.x; x.v[5] = hi.y; x.v[6] = hi.z; x.v[7] = hi.w;
  return x;
}
// stage twiddles: dense per-stage tables, entry (b, i) = w_{2^{b+1}}^i at offset 2^b - 1 + i
__device__ __forceinline__ Fr tw_load(const uint4* tw, int b, int i) {
  int o = (1 << b) - 1 + i;
  uint4 lo = tw[o], hi = tw[kStageTw + o];
  Fr x;
  x.v[0] = lo.x; x.v[1] = lo.y; x.v[2] = lo.z; x.v[3] = lo.w;
  x.v[4] = hi.x; x.v[5] = hi.y; x.v[6] = hi.z; x.v[7] = hi.w;
  return x;
}

__device__ __forceinline__ bool fr_is_canonical(const Fr& x) {
  bool lt = false, decided = false;
#pragma unroll
  for (int k = 7; k >= 0; k--) {
    if (!decided && x.v[k] != FrParams::mod(k)) {
      lt = x.v[k] < FrParams::mod(k);
      decided = true;
    }
  }
  return lt;
}

// decimation-in-frequency butterfly
__device__ __forceinline__ void bfly(Fr& a, Fr& b, const Fr& w) {
  Fr t = a - b;
  a = a + b;
  b = t * w;
}
__device__ __forceinline__ void bfly_trivial(Fr& a, Fr& b) {
  Fr t = a - b;
  a = a + b;
  b = t;
}

__device__ __forceinline__ uint32_t bitrev(uint32_t x, int bits) {
  return bits ? (__brev(x) >> (32 - bits)) : 0;
}

// digit reversal of a row number: digits (widths w[0..nd-1]) are stored most-significant first
// in `rho`; the output puts digit 0 in the least-significant position.
__device__ __forceinline__ uint32_t row_reverse(uint32_t rho, const NttPassArgs& p) {
  uint32_t out = 0;
  int shift_in = p.row_bits;
  int shift_out = 0;
#pragma unroll 1
  for (int d = 0; d < p.n_row_digits; d++) {
    int w = p.row_digit_bits[d];
    shift_in -= w;
    uint32_t dig = (rho >> shift_in) & ((1u << w) - 1);
    out |= dig << shift_out;
    shift_out += w;
  }
  return out;
}

template <bool FINAL>
__global__ void __launch_bounds__(kThreads, 2) ntt_pass_kernel(NttPassArgs p) {
  extern __shared__ uint4 smem[];
  uint4* data = smem;
  uint4* stw = smem + kDataU4;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int k = p.k;
  const int gbits = 8 - k;                    // sub-columns per warp = 2^gbits
  const int ct_bits = 3 + gbits;              // columns per tile = 8 << gbits
  const int CT = 1 << ct_bits;

  // stage twiddles -> shared (two planes)
  {
    const uint4* g = reinterpret_cast<const uint4*>(p.stage_tw);
    for (int i = tid; i < 255; i += kThreads) {
      stw[i] = g[2 * i];
      stw[kStageTw + i] = g[2 * i + 1];
    }
  }

  // ---- tile coordinates
  uint32_t tile = blockIdx.x;
  uint64_t base = 0;      // column mode: first element of (blk, r_base)
  uint32_t r_base = 0;    // column mode
  uint32_t c_base = 0;    // final mode: first K_0 of this tile
  uint32_t mid = 0;       // final mode: the remaining (lower) row digits
  uint64_t batch_off = 0; // final mode: first element of this tile's transform within a batch
  uint32_t ncols;         // valid columns in this tile
  if (!FINAL) {
    uint32_t tiles_per_blk = (p.S + CT - 1) >> ct_bits;
    uint32_t blk = tile / tiles_per_blk;
    r_base = (tile - blk * tiles_per_blk) << ct_bits;
    base = uint64_t(blk) * p.M + r_base;
    ncols = min(uint32_t(CT), p.S - r_base);
  } else if (p.batch_cols) {
    c_base = tile << ct_bits;               // first transform of this tile
    ncols = min(uint32_t(CT), p.batch_count - c_base);
  } else {
    // a contiguous batch of transforms: tile -> (transform, tile within it)
    if (p.tiles_per_batch) {
      uint32_t bidx = tile / p.tiles_per_batch;
      tile -= bidx * p.tiles_per_batch;
      batch_off = uint64_t(bidx) << p.batch_log;
    }
    // row number rho = K0 * 2^{mid_bits} + mid ; tiles enumerate (mid, K0 / CT)
    uint32_t k0_range = 1u << p.first_digit_bits;
    uint32_t tiles_per_mid = (k0_range + CT - 1) >> ct_bits;
    mid = tile / tiles_per_mid;
    c_base = (tile - mid * tiles_per_mid) << ct_bits;
    ncols = min(uint32_t(CT), k0_range - c_base);
  }
  const int mid_bits = p.row_bits - p.first_digit_bits;

  // ---- stage in: global -> shared, one whole element per thread per step
  const char* in = reinterpret_cast<const char*>(p.in);
#pragma unroll 2
  for (int it = 0; it < kTileElems / kThreads; it++) {
    uint32_t q = it * kThreads + tid;
    uint32_t c = q & (CT - 1);
    uint32_t t = q >> ct_bits;
    if (FINAL && p.batch_cols && !p.in_ilv) {   // contiguous transforms: consecutive threads read one of them
      t = q & ((1u << k) - 1);
      c = q >> k;
    }
    Fr x = Fr::zero();
    if (c < ncols) {
      uint64_t idx;
      if (!FINAL) idx = base + uint64_t(t) * p.S + c;
      else if (p.batch_cols) idx = (uint64_t(c_base + c) << k) + t;
      else idx = batch_off + ((uint64_t(c_base + c) << mid_bits) + mid) * (1u << k) + t;
      uint64_t src = idx;
      if (p.in_ilv) src = ((idx & ((1ull << p.xform_log) - 1)) << p.in_ilv_log) + (idx >> p.xform_log);
      x = Fr::load_nc(in + src * 32);
      if (!p.mont_in) {
        if (!fr_is_canonical(x)) atomicOr(p.flag, 1u);
        x = x.to_mont();
      }
      if (p.pre_scale) x = x * Fr::load_nc(reinterpret_cast<const char*>(p.pre_scale) + idx * 32);
    }
    sm_store(data, c >> gbits, ((c & ((1 << gbits) - 1)) << k) | t, x);
  }
  __syncthreads();

  // ---- per-warp transforms; slot u = g * 2^k + t, stage b acts on bit b of u (b < k)
  Fr x[8];
  if (k > 5) {
    // group A: slots j <-> bits 7..5 ; u = j*32 + lane
#pragma unroll
    for (int j = 0; j < 8; j++) x[j] = sm_load(data, warp, j * 32 + lane);
    if (k > 7) {
#pragma unroll
      for (int j = 0; j < 4; j++) bfly(x[j], x[j + 4], tw_load(stw, 7, j * 32 + lane));
    }
    if (k > 6) {
#pragma unroll
      for (int j = 0; j < 8; j++)
        if (!(j & 2)) bfly(x[j], x[j + 2], tw_load(stw, 6, (j & 1) * 32 + lane));
    }
    {
      Fr w = tw_load(stw, 5, lane);
#pragma unroll
      for (int j = 0; j < 8; j += 2) bfly(x[j], x[j + 1], w);
    }
#pragma unroll
    for (int j = 0; j < 8; j++) sm_store(data, warp, j * 32 + lane, x[j]);
    __syncwarp();
  }
  if (k > 2) {
    // group B: slots j <-> bits 4..2 ; u = hi*32 + j*4 + lo, lane = hi*4 + lo
    const int hi = lane >> 2, lo = lane & 3;
#pragma unroll
    for (int j = 0; j < 8; j++) x[j] = sm_load(data, warp, hi * 32 + j * 4 + lo);
    if (k > 4) {
#pragma unroll
      for (int j = 0; j < 4; j++) bfly(x[j], x[j + 4], tw_load(stw, 4, j * 4 + lo));
    }
    if (k > 3) {
#pragma unroll
      for (int j = 0; j < 8; j++)
        if (!(j & 2)) bfly(x[j], x[j + 2], tw_load(stw, 3, (j & 1) * 4 + lo));
    }
    {
      Fr w = tw_load(stw, 2, lo);
#pragma unroll
      for (int j = 0; j < 8; j += 2) bfly(x[j], x[j + 1], w);
    }
#pragma unroll
    for (int j = 0; j < 8; j++) sm_store(data, warp, hi * 32 + j * 4 + lo, x[j]);
    __syncwarp();
  }
  {
    // group C: slots j <-> bits 2..0 ; u = lane*8 + j ; stages b = 1 (w_4^{j&1}) and b = 0
#pragma unroll
    for (int j = 0; j < 8; j++) x[j] = sm_load(data, warp, lane * 8 + j);
    if (k > 1) {
      Fr w4 = tw_load(stw, 1, 1);
      bfly_trivial(x[0], x[2]);
      bfly(x[1], x[3], w4);
      bfly_trivial(x[4], x[6]);
      bfly(x[5], x[7], w4);
    }
#pragma unroll
    for (int j = 0; j < 8; j += 2) bfly_trivial(x[j], x[j + 1]);
    // store in natural frequency order: slot (g, t) -> (g, bitrev_k(t))
    const uint32_t tmask = (1u << k) - 1;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      uint32_t u = lane * 8 + j;
      uint32_t v = (u & ~tmask) | bitrev(u & tmask, k);
      sm_store(data, warp, v, x[j]);
    }
  }
  __syncthreads();

  // ---- stage out: shared -> global (+ inter-pass twiddle / post scale / representation change)
  char* out = reinterpret_cast<char*>(p.out);
#pragma unroll 2
  for (int it = 0; it < kTileElems / kThreads; it++) {
    uint32_t q = it * kThreads + tid;
    uint32_t c = q & (CT - 1);
    uint32_t K = q >> ct_bits;
    if (FINAL && p.batch_cols) {        // consecutive threads walk one transform: contiguous 32 B * 2^k runs
      K = q & ((1u << k) - 1);
      c = q >> k;
    }
    if (c >= ncols) continue;
    Fr y = sm_load(data, c >> gbits, ((c & ((1 << gbits) - 1)) << k) | K);
    uint64_t idx;
    if (!FINAL) {
      idx = base + uint64_t(K) * p.S + c;
      if (p.tw_full) {
        // table indexed by the position inside the size-M block
        uint64_t off = uint64_t(K) * p.S + r_base + c;
        y = y * Fr::load_nc(reinterpret_cast<const char*>(p.tw_full) + off * 32);
      } else {
        // w_M^{r K} = w_n^{r K (n/M)} from the two-level table
        uint64_t e = (uint64_t(r_base + c) * K) << p.tw_shift;
        Fr w = Fr::load_nc(reinterpret_cast<const char*>(p.tw_lo) + (e & p.tw_lo_mask) * 32) *
               Fr::load_nc(reinterpret_cast<const char*>(p.tw_hi) + (e >> p.tw_lo_bits) * 32);
        y = y * w;
      }
    } else {
      if (p.batch_cols) {
        idx = (uint64_t(c_base + c) << k) + K;
      } else {
        uint32_t rho = ((c_base + c) << mid_bits) + mid;
        idx = batch_off + uint64_t(row_reverse(rho, p)) + (uint64_t(K) << p.row_bits);
      }
      if (p.post_scale)
        y = y * Fr::load_nc(reinterpret_cast<const char*>(p.post_scale) + idx * 32);
    }
    if (!p.mont_out) y = y.from_mont();
    if (FINAL && p.out_win) {
      const uint64_t b = idx >> p.xform_log, i = idx & ((1ull << p.xform_log) - 1);
      const uint64_t col = i & ((1ull << p.out_col_bits) - 1);
      char* dst = reinterpret_cast<char*>(p.out_wins[i >> p.out_col_bits]);
      y.store(dst + ((b << p.out_col_bits) + col) * 32);
    } else {
      y.store(out + idx * 32);
    }
  }
}

// out[i] = c * base^i   (Montgomery), one thread per entry, square-and-multiply on the index
__global__ void fr_powers_kernel(Fr* out, Fr base, Fr c, uint64_t count, uint64_t stride_pow) {
  uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
  if (i >= count) return;
  uint64_t e = i * stride_pow;
  Fr r = c, b = base;
  while (e) {
    if (e & 1) r = r * b;
    b = b.sqr();
    e >>= 1;
  }
  r.store(out + i);
}

// full inter-pass twiddle table for one column pass: tab[K*S + r] = w_M^{r*K}
__global__ void fr_twiddle_full_kernel(Fr* tab, const Fr* lo, const Fr* hi, uint32_t lo_bits,
                                       uint32_t tw_shift, uint32_t s_bits, uint64_t M) {
  uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
  if (i >= M) return;
  uint64_t K = i >> s_bits, r = i & ((1ull << s_bits) - 1);
  uint64_t e = (r * K) << tw_shift;
  Fr w = Fr::load(lo + (e & ((1ull << lo_bits) - 1))) * Fr::load(hi + (e >> lo_bits));
  w.store(tab + i);
}

// element-wise helpers used by the prover pipeline
__global__ void fr_convert_kernel(Fr* data, uint64_t n, int to_mont) {
  uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  Fr x = Fr::load(data + i);
  x = to_mont ? x.to_mont() : x.from_mont();
  x.store(data + i);
}

// layout-d block of rank g out of a natural-order vector (csrc/dist.cu):
// out[low_loc * H + high] = full[(g * lows + low_loc) + (high << d)],  H = n >> d
__global__ void fr_gather_layout_kernel(const uint4* __restrict__ full, uint4* __restrict__ out,
                                        uint32_t log_n, uint32_t d, uint32_t lows, uint32_t low0) {
  uint64_t q = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
  const uint32_t hbits = log_n - d;
  if (q >= (uint64_t(lows) << hbits) * 2) return;
  uint64_t pos = q >> 1;
  uint64_t low_loc = pos >> hbits, high = pos & ((1ull << hbits) - 1);
  uint64_t src = (low0 + low_loc) + (high << d);
  out[q] = full[(src << 1) | (q & 1)];
}

__global__ void fr_table_kernel(Fr* out, Fr base, Fr c, uint64_t rows, uint64_t cols, uint64_t row0,
                                uint64_t mul, int mode) {
  uint64_t pos = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
  if (pos >= rows * cols) return;
  uint64_t hi = pos / cols, lo = pos - hi * cols;
  uint64_t e = mode == 0 ? (row0 + hi) * lo : (row0 + hi) + lo * mul;
  Fr r = c, b = base;
  while (e) {
    if (e & 1) r = r * b;
    b = b.sqr();
    e >>= 1;
  }
  r.store(out + pos);
}

__global__ void fr_mul_sub_kernel(Fr* a, const Fr* b, const Fr* c, uint64_t n) {
  uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  (Fr::load(a + i) * Fr::load(b + i) - Fr::load(c + i)).store(a + i);
}

// ---------------------------------------------------------------- host side
void ntt_launch_pass(const NttPassArgs& a, bool final, uint32_t tiles, cudaStream_t s) {
  // the opt-in to 64 KB of dynamic shared memory is per device: once for every GPU this process uses
  // (several contexts / host threads may launch concurrently)
  static std::once_flag configured[64];
  int dev = 0;
  cudaGetDevice(&dev);
  std::call_once(configured[dev & 63], [] {
    cudaFuncSetAttribute(ntt_pass_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         int(kSmemBytes));
    cudaFuncSetAttribute(ntt_pass_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         int(kSmemBytes));
  });
  if (final)
    ntt_pass_kernel<true><<<tiles, kThreads, kSmemBytes, s>>>(a);
  else
    ntt_pass_kernel<false><<<tiles, kThreads, kSmemBytes, s>>>(a);
  count_launches(1);
}

void fr_powers_launch(void* out, const uint32_t base[8], const uint32_t c[8], uint64_t count,
                      uint64_t stride_pow, cudaStream_t s) {
  Fr b, cc;
  for (int i = 0; i < 8; i++) { b.v[i] = base[i]; cc.v[i] = c[i]; }
  uint32_t blocks = uint32_t((count + 255) / 256);
  if (blocks) fr_powers_kernel<<<blocks, 256, 0, s>>>(reinterpret_cast<Fr*>(out), b, cc, count, stride_pow);
  count_launches(blocks ? 1 : 0);
}

void fr_twiddle_full_launch(void* tab, const void* lo, const void* hi, uint32_t lo_bits,
                            uint32_t tw_shift, uint32_t s_bits, uint64_t M, cudaStream_t s) {
  uint32_t blocks = uint32_t((M + 255) / 256);
  fr_twiddle_full_kernel<<<blocks, 256, 0, s>>>(reinterpret_cast<Fr*>(tab),
                                                reinterpret_cast<const Fr*>(lo),
                                                reinterpret_cast<const Fr*>(hi), lo_bits, tw_shift,
                                                s_bits, M);
  count_launches(blocks ? 1 : 0);
}

void fr_gather_layout_launch(const void* full, void* out, uint32_t log_n, uint32_t d, uint32_t world,
                             uint32_t rank, cudaStream_t s) {
  uint32_t lows = (1u << d) / world;
  uint64_t total = (uint64_t(lows) << (log_n - d)) * 2;
  uint32_t blocks = uint32_t((total + 255) / 256);
  if (blocks)
    fr_gather_layout_kernel<<<blocks, 256, 0, s>>>(reinterpret_cast<const uint4*>(full),
                                                   reinterpret_cast<uint4*>(out), log_n, d, lows, rank * lows);
  count_launches(blocks ? 1 : 0);
}

void fr_table_launch(void* out, const uint32_t base[8], const uint32_t c[8], uint64_t rows, uint64_t cols,
                     uint64_t row0, uint64_t mul, int mode, cudaStream_t s) {
  Fr b, cc;
  for (int i = 0; i < 8; i++) { b.v[i] = base[i]; cc.v[i] = c[i]; }
  uint32_t blocks = uint32_t((rows * cols + 255) / 256);
  if (blocks) fr_table_kernel<<<blocks, 256, 0, s>>>(reinterpret_cast<Fr*>(out), b, cc, rows, cols, row0, mul, mode);
  count_launches(blocks ? 1 : 0);
}

void fr_mul_sub_launch(void* a, const void* b, const void* c, uint64_t n, cudaStream_t s) {
  uint32_t blocks = uint32_t((n + 255) / 256);
  if (blocks)
    fr_mul_sub_kernel<<<blocks, 256, 0, s>>>(reinterpret_cast<Fr*>(a), reinterpret_cast<const Fr*>(b),
                                             reinterpret_cast<const Fr*>(c), n);
  count_launches(blocks ? 1 : 0);
}

void fr_convert_launch(void* data, uint64_t n, int to_mont, cudaStream_t s) {
  uint32_t blocks = uint32_t((n + 255) / 256);
  if (blocks) fr_convert_kernel<<<blocks, 256, 0, s>>>(reinterpret_cast<Fr*>(data), n, to_mont);
  count_launches(blocks ? 1 : 0);
}

}  // namespace zkb
