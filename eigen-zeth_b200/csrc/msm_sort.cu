// msm_sort.cu — K5: scalars -> (bucket key, point index) pairs grouped by bucket, hand-written.
//
// No reference counterpart.  The walk (msm_impl.cuh) only needs entries of one
// bucket to be CONTIGUOUS and buckets to appear in key order; it does not need a stable sort.  That
// allows a two-level MSD bucket sort fused with the digit extraction, instead of digits -> 16 B per
// entry in HBM -> a generic 3-pass LSD radix sort:
//
//   A  histogram   read the scalars, recompute every signed digit, count entries per COARSE bin
//                  (key >> lo_bits, at most 2^11 bins) in shared memory, one global atomic per (tile, bin).
//   P  plan        one CTA: exclusive scan of the coarse histogram (= where each bin starts in the
//                  output AND in the intermediate array), m_valid, and the work list of pass C:
//                  one item per bin, or several per bin when a bin is larger than a CTA can hold.
//   B  scatter     a tile = the <= kTile entries of ~kTile / windows scalars.  The tile is counting-sorted
//                  by coarse bin INSIDE shared memory (count, scan, place), room is reserved per
//                  (tile, bin) with one global atomic, and the staged tile is written out linearly:
//                  neighbouring threads write neighbouring entries of one run (~5 entries = 40 B per
//                  bin per tile with 2^11 bins), so the stores coalesce and every run is written at
//                  once.  Two tiles are resident per SM so that one tile's barriers hide under the
//                  other's work (measured at 2^24: 10 K-entry tiles, two per SM, 0.98 ms; 20 K-entry
//                  tiles, one per SM, 1.09 ms; 512-thread CTAs with 4-8 K tiles, 1.3-1.4 ms).  (A first
//                  version scattered straight from registers: each of the ~450 resident tiles then kept
//                  16 K half-written lines open, 900 MB of them — far beyond the 126 MB L2 — and the
//                  pass took 7 ms at 2^24 points.)
//   C  bucketize   one CTA per work item loads its segment (<= kSeg entries, 8 B each) into shared
//                  memory, counting-sorts it by fine bucket (key & (2^lo_bits - 1)) in shared memory
//                  and writes it out linearly to the bucket's final position.  A bin that fits one
//                  segment is the CTA's private output region (offsets from a local scan).  Larger
//                  bins — the rule at 2^24 points, and any skewed input (0/1 witnesses, all-equal
//                  scalars) — are cut into segments that first add their fine histograms into a global
//                  table (C1), get bucket offsets from a per-bin scan (C1s) and reserve their output
//                  runs with global atomics (C2): same result for any distribution.
// Zero digits are never emitted (no sentinel key, no count kernel).  HBM traffic per entry: 8 B
// written + 8 B read (+ 8 B read for C1) + 8 B written, plus 2 x 32 B per scalar — against 16 B +
// 3 x 16 B + histogram for digits + cub::DeviceRadixSort.  The order inside a bucket is not
// deterministic; the bucket SUM is, and results are compared in canonical affine form.
#include <cuda_runtime.h>

#include <mutex>

#include "msm.h"

namespace zkb {

constexpr uint32_t kSortThreads = 512;
constexpr uint32_t kPlaceThreads = 1024;
constexpr uint32_t kCap = 20480;            // 2 * kSeg: a coarse bin up to kCap / 6 entries on average stays private
constexpr uint32_t kTile = kCap / 2;        // entries of a pass-B tile (80 KB staged + 16 KB bin tables: two CTAs per SM)
constexpr uint32_t kMaxCoarse = 2048;       // coarse bins: ~10 entries per (tile, bin) run
constexpr uint32_t kMaxFine = 4096;         // fine buckets per coarse bin

// signed-digit extraction shared by passes A and B (same recoding as msm_digits_kernel)
struct DigitCfg {
  uint32_t c, nwin, half, mask, val_stride, idx_base, one_set;
};

__device__ __forceinline__ bool scalar_is_canonical(const uint32_t (&s)[9]) {
  const uint32_t rmod[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u,
                            0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
  bool lt = false, decided = false;
#pragma unroll
  for (int k = 7; k >= 0; k--) {
    if (!decided && s[k] != rmod[k]) {
      lt = s[k] < rmod[k];
      decided = true;
    }
  }
  return lt;
}

// calls f(key, val) for every non-zero digit of scalar i.  C > 0 fixes the window width at compile
// time: the loop unrolls, every window's word index and shift become immediates and the scalar's
// limbs stay in registers (with a run-time width the limb array is indexed dynamically and lives in
// local memory — the digit loop, not HBM, then bounds passes A and B).  C == 0 is the generic path.
template <uint32_t C, class Fn>
__device__ __forceinline__ void for_each_digit(const uint32_t* __restrict__ scalars, uint64_t i, const DigitCfg& d,
                                               uint32_t* flag, bool check, Fn f) {
  const uint4* sp = reinterpret_cast<const uint4*>(scalars + 8 * i);
  uint4 lo = __ldg(sp), hi = __ldg(sp + 1);
  uint32_t s[9] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w, 0};
  if (check && !scalar_is_canonical(s)) atomicOr(flag, 1u);
  uint32_t carry = 0;
  const uint32_t idx = d.idx_base + uint32_t(i);
  if constexpr (C != 0) {
    constexpr uint32_t kWin = (255 + C - 1) / C, kMask = (1u << C) - 1, kHalf = 1u << (C - 1);
#pragma unroll
    for (uint32_t w = 0; w < kWin; w++) {
      const uint32_t bit = w * C, word = bit >> 5, sh = bit & 31;
      uint32_t dg = (word < 8 ? __funnelshift_r(s[word], s[word + 1], sh) & kMask : 0u) + carry;
      uint32_t neg = dg > kHalf;
      uint32_t mag = neg ? (1u << C) - dg : dg;
      carry = neg;
      if (mag) f((d.one_set ? 0u : w * kHalf) + (mag - 1), (w * d.val_stride + idx) | (neg << 31));
    }
  } else {
    for (uint32_t w = 0; w < d.nwin; w++) {
      uint32_t bit = w * d.c;
      uint32_t word = bit >> 5, sh = bit & 31;
      uint32_t dg = 0;
      if (word < 8) dg = __funnelshift_r(s[word], s[word + 1], sh) & d.mask;
      dg += carry;
      uint32_t neg = dg > d.half;
      uint32_t mag = neg ? (1u << d.c) - dg : dg;
      carry = neg;
      if (mag) f((d.one_set ? 0u : w * d.half) + (mag - 1), (w * d.val_stride + idx) | (neg << 31));
    }
  }
}

// ---------------------------------------------------------------- A: coarse histogram
template <uint32_t C>
__global__ void __launch_bounds__(kSortThreads) sort_hist_kernel(const uint32_t* __restrict__ scalars, uint64_t n,
                                                                 DigitCfg d, uint32_t lo_bits, uint32_t n_coarse,
                                                                 uint32_t per_cta, uint32_t* __restrict__ coarse_hist,
                                                                 uint32_t* __restrict__ flag) {
  extern __shared__ uint32_t sm_cnt[];
  for (uint32_t b = threadIdx.x; b < n_coarse; b += blockDim.x) sm_cnt[b] = 0;
  __syncthreads();
  const uint64_t first = uint64_t(blockIdx.x) * per_cta;
  const uint64_t last = first + per_cta < n ? first + per_cta : n;
  for (uint64_t i = first + threadIdx.x; i < last; i += blockDim.x)
    for_each_digit<C>(scalars, i, d, flag, true, [&](uint32_t key, uint32_t) { atomicAdd(&sm_cnt[key >> lo_bits], 1u); });
  __syncthreads();
  for (uint32_t b = threadIdx.x; b < n_coarse; b += blockDim.x) {
    uint32_t v = sm_cnt[b];
    if (v) atomicAdd(&coarse_hist[b], v);
  }
}

// ---------------------------------------------------------------- P: plan (one CTA)
// in: coarse_hist[n_coarse].  out: coarse_off[n_coarse + 1] (exclusive scan), cursor[b] = coarse_off[b]
// (pass B's write frontiers), m_valid, items[] = (bin, segment) for pass C, n_items, and
// bin_flags[b] = 1 for bins cut into several segments.
__global__ void __launch_bounds__(1024) sort_plan_kernel(const uint32_t* __restrict__ coarse_hist, uint32_t n_coarse,
                                                         uint32_t* __restrict__ coarse_off, uint32_t* __restrict__ cursor,
                                                         uint64_t* __restrict__ m_valid, uint2* __restrict__ items,
                                                         uint32_t* __restrict__ n_items, uint32_t* __restrict__ bin_multi) {
  __shared__ uint32_t s_sum[1024], s_seg[1024];
  const uint32_t t = threadIdx.x, per = (n_coarse + 1023) / 1024;
  const uint32_t lo = t * per, hi = lo + per < n_coarse ? lo + per : n_coarse;
  uint32_t sum = 0, segs = 0;
  for (uint32_t b = lo; b < hi; b++) {
    uint32_t v = coarse_hist[b];
    sum += v;
    segs += (v + kCap / 2 - 1) / (kCap / 2);
  }
  s_sum[t] = sum;
  s_seg[t] = segs;
  __syncthreads();
  // Hillis-Steele inclusive scan over 1024 partials (both arrays)
  for (uint32_t o = 1; o < 1024; o <<= 1) {
    uint32_t a = t >= o ? s_sum[t - o] : 0, b2 = t >= o ? s_seg[t - o] : 0;
    __syncthreads();
    s_sum[t] += a;
    s_seg[t] += b2;
    __syncthreads();
  }
  uint32_t off = s_sum[t] - sum, item = s_seg[t] - segs;
  for (uint32_t b = lo; b < hi; b++) {
    uint32_t v = coarse_hist[b];
    coarse_off[b] = off;
    cursor[b] = off;
    uint32_t ns = (v + kCap / 2 - 1) / (kCap / 2);
    bin_multi[b] = ns > 1;
    for (uint32_t s = 0; s < ns; s++) items[item + s] = make_uint2(b, s);
    item += ns;
    off += v;
  }
  if (t == 1023) {
    coarse_off[n_coarse] = s_sum[1023];
    *m_valid = s_sum[1023];
    *n_items = s_seg[1023];
  }
}

// exclusive scan of v[0..n) in shared memory by the whole CTA (blockDim.x a multiple of 32, <= 1024;
// n <= 4 * blockDim.x is typical); `part` is >= 32 words of scratch.  On return v[i] = sum of the old
// v[0..i).  Warp shuffles + one pass over the warp totals: three barriers, not two per doubling step.
__device__ __forceinline__ void block_exclusive_scan(uint32_t* v, uint32_t n, uint32_t* part) {
  const uint32_t t = threadIdx.x, T = blockDim.x, per = (n + T - 1) / T, lane = t & 31, warp = t >> 5;
  const uint32_t lo = t * per, hi = lo + per < n ? lo + per : n;
  uint32_t sum = 0;
  for (uint32_t i = lo; i < hi; i++) sum += v[i];
  uint32_t inc = sum;
#pragma unroll
  for (uint32_t o = 1; o < 32; o <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += y;
  }
  if (lane == 31) part[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const uint32_t w = lane < (T >> 5) ? part[lane] : 0;
    uint32_t winc = w;
#pragma unroll
    for (uint32_t o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += y;
    }
    part[lane] = winc - w;
  }
  __syncthreads();
  uint32_t off = part[warp] + inc - sum;
  for (uint32_t i = lo; i < hi; i++) {
    uint32_t c = v[i];
    v[i] = off;
    off += c;
  }
  __syncthreads();
}

// ---------------------------------------------------------------- B: scatter into coarse bins
// dynamic shared memory: kCap staged entries | n_coarse run cursors | n_coarse (global base - local start)
template <uint32_t C>
__global__ void __launch_bounds__(kPlaceThreads, 2) sort_scatter_kernel(const uint32_t* __restrict__ scalars, uint64_t n,
                                                                     DigitCfg d, uint32_t lo_bits, uint32_t n_coarse,
                                                                     uint32_t per_cta, uint32_t* __restrict__ cursor,
                                                                     uint2* __restrict__ mid) {
  extern __shared__ uint2 sm_stage[];
  uint32_t* run = reinterpret_cast<uint32_t*>(sm_stage + kTile);
  uint32_t* gbase = run + n_coarse;
  __shared__ uint32_t part[32];
  __shared__ uint32_t total;
  const uint32_t t = threadIdx.x;
  for (uint32_t b = t; b < n_coarse; b += blockDim.x) run[b] = 0;
  __syncthreads();
  const uint64_t first = uint64_t(blockIdx.x) * per_cta;
  const uint64_t last = first + per_cta < n ? first + per_cta : n;
  for (uint64_t i = first + t; i < last; i += blockDim.x)
    for_each_digit<C>(scalars, i, d, nullptr, false, [&](uint32_t key, uint32_t) { atomicAdd(&run[key >> lo_bits], 1u); });
  __syncthreads();
  // reserve this tile's run in every bin it touches, then turn the counts into local start offsets
  for (uint32_t b = t; b < n_coarse; b += blockDim.x) {
    uint32_t v = run[b];
    gbase[b] = v ? atomicAdd(&cursor[b], v) : 0;
  }
  __syncthreads();
  block_exclusive_scan(run, n_coarse, part);
  for (uint32_t b = t; b < n_coarse; b += blockDim.x) gbase[b] -= run[b];     // global pos = gbase[bin] + local index
  __syncthreads();
  for (uint64_t i = first + t; i < last; i += blockDim.x)
    for_each_digit<C>(scalars, i, d, nullptr, false, [&](uint32_t key, uint32_t val) {
      uint32_t pos = atomicAdd(&run[key >> lo_bits], 1u);
      sm_stage[pos] = make_uint2(key, val);
    });
  __syncthreads();
  if (t == 0) total = run[n_coarse - 1];          // the last bin's cursor ends at the tile's entry count
  __syncthreads();
  for (uint32_t i = t; i < total; i += blockDim.x) {
    uint2 e = sm_stage[i];
    mid[gbase[e.x >> lo_bits] + i] = e;
  }
}

// visits mid[s0 + i] for i in [0, cnt) with the whole CTA, four independent loads in flight per
// thread before the dependent shared-memory atomics (these loops are latency-bound otherwise)
template <class Fn>
__device__ __forceinline__ void for_each_entry(const uint2* __restrict__ mid, uint32_t s0, uint32_t cnt, Fn f) {
  const uint32_t t = threadIdx.x, T = blockDim.x;
  for (uint32_t i = t; i < cnt; i += 4 * T) {
    uint2 e[4];
#pragma unroll
    for (uint32_t j = 0; j < 4; j++)
      if (i + j * T < cnt) e[j] = mid[s0 + i + j * T];
#pragma unroll
    for (uint32_t j = 0; j < 4; j++)
      if (i + j * T < cnt) f(e[j]);
  }
}

// ---------------------------------------------------------------- C1: fine histograms of oversized bins
__global__ void __launch_bounds__(kSortThreads) sort_fine_hist_kernel(const uint2* __restrict__ mid,
                                                                      const uint32_t* __restrict__ coarse_off,
                                                                      const uint2* __restrict__ items,
                                                                      const uint32_t* __restrict__ n_items,
                                                                      const uint32_t* __restrict__ bin_multi,
                                                                      uint32_t lo_bits, uint32_t* __restrict__ fine) {
  __shared__ uint32_t cnt[kMaxFine];
  // a CTA takes a CONTIGUOUS range of work items: the segments of one bin are neighbours in the list,
  // so the shared-memory histogram is flushed (with atomics: a bin can straddle two CTAs) only when
  // the bin changes, not once per segment
  const uint32_t total = *n_items, per = (total + gridDim.x - 1) / gridDim.x;
  const uint32_t it0 = blockIdx.x * per, it1 = it0 + per < total ? it0 + per : total;
  const uint32_t L = 1u << lo_bits, mask = L - 1;
  uint32_t open_bin = 0xffffffffu;
  for (uint32_t it = it0; it < it1; it++) {
    const uint2 w = items[it];
    if (w.x != open_bin) {
      if (open_bin != 0xffffffffu) {
        __syncthreads();
        for (uint32_t f = threadIdx.x; f < L; f += blockDim.x) {
          uint32_t v = cnt[f];
          if (v) atomicAdd(&fine[(uint64_t(open_bin) << lo_bits) + f], v);
        }
        open_bin = 0xffffffffu;
      }
      if (!bin_multi[w.x]) continue;
      __syncthreads();
      for (uint32_t f = threadIdx.x; f < L; f += blockDim.x) cnt[f] = 0;
      __syncthreads();
      open_bin = w.x;
    }
    const uint32_t b0 = coarse_off[w.x], b1 = coarse_off[w.x + 1];
    const uint32_t s0 = b0 + w.y * (kCap / 2), s1 = s0 + kCap / 2 < b1 ? s0 + kCap / 2 : b1;
    for_each_entry(mid, s0, s1 - s0, [&](uint2 e) { atomicAdd(&cnt[e.x & mask], 1u); });
  }
  if (open_bin != 0xffffffffu) {
    __syncthreads();
    for (uint32_t f = threadIdx.x; f < L; f += blockDim.x) {
      uint32_t v = cnt[f];
      if (v) atomicAdd(&fine[(uint64_t(open_bin) << lo_bits) + f], v);
    }
  }
}

// C1s: per oversized bin, fine counts -> output cursors (exclusive scan + the bin's base)
__global__ void __launch_bounds__(1024) sort_fine_scan_kernel(const uint32_t* __restrict__ coarse_off,
                                                              const uint32_t* __restrict__ bin_multi, uint32_t n_coarse,
                                                              uint32_t lo_bits, uint32_t* __restrict__ fine) {
  __shared__ uint32_t part[1024];
  for (uint32_t b = blockIdx.x; b < n_coarse; b += gridDim.x) {
    if (!bin_multi[b]) continue;
    const uint32_t L = 1u << lo_bits, per = (L + 1023) / 1024, t = threadIdx.x;
    uint32_t* f = fine + (uint64_t(b) << lo_bits);
    const uint32_t lo = t * per, hi = lo + per < L ? lo + per : L;
    uint32_t sum = 0;
    for (uint32_t i = lo; i < hi; i++) sum += f[i];
    part[t] = sum;
    __syncthreads();
    for (uint32_t o = 1; o < 1024; o <<= 1) {
      uint32_t a = t >= o ? part[t - o] : 0;
      __syncthreads();
      part[t] += a;
      __syncthreads();
    }
    uint32_t off = coarse_off[b] + part[t] - sum;
    for (uint32_t i = lo; i < hi; i++) {
      uint32_t v = f[i];
      f[i] = off;
      off += v;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------- C2: final placement
// dynamic shared memory: kSeg sorted entries | 2^lo_bits run cursors | 2^lo_bits (global base - local
// start).  The segment is read twice from global memory (count, then place: the second read hits L2)
// rather than staged, so that two CTAs fit on an SM.
constexpr uint32_t kSeg = kCap / 2;
__global__ void __launch_bounds__(kPlaceThreads, 2) sort_place_kernel(const uint2* __restrict__ mid,
                                                                  const uint32_t* __restrict__ coarse_off,
                                                                  const uint2* __restrict__ items,
                                                                  const uint32_t* __restrict__ n_items,
                                                                  const uint32_t* __restrict__ bin_multi, uint32_t lo_bits,
                                                                  uint32_t* __restrict__ fine, uint32_t* __restrict__ keys_out,
                                                                  uint32_t* __restrict__ vals_out) {
  extern __shared__ uint2 sorted[];
  uint32_t* run = reinterpret_cast<uint32_t*>(sorted + kSeg);
  uint32_t* gbase = run + (1u << lo_bits);
  __shared__ uint32_t part[32];
  for (uint32_t it = blockIdx.x; it < *n_items; it += gridDim.x) {
    const uint2 w = items[it];
    const uint32_t L = 1u << lo_bits, mask = L - 1, t = threadIdx.x;
    for (uint32_t f = t; f < L; f += blockDim.x) run[f] = 0;
    __syncthreads();
    const uint32_t b0 = coarse_off[w.x], b1 = coarse_off[w.x + 1];
    const uint32_t s0 = b0 + w.y * kSeg, s1 = s0 + kSeg < b1 ? s0 + kSeg : b1, cnt = s1 - s0;
    for_each_entry(mid, s0, cnt, [&](uint2 e) { atomicAdd(&run[e.x & mask], 1u); });
    __syncthreads();
    if (bin_multi[w.x]) {
      // shared bin: reserve this segment's run in every fine bucket from the global cursors
      for (uint32_t f = t; f < L; f += blockDim.x) {
        uint32_t v = run[f];
        gbase[f] = v ? atomicAdd(&fine[(uint64_t(w.x) << lo_bits) + f], v) : 0;
      }
      __syncthreads();
      block_exclusive_scan(run, L, part);
      for (uint32_t f = t; f < L; f += blockDim.x) gbase[f] -= run[f];
    } else {
      // private bin: the local scan IS the layout of the bin's output region
      block_exclusive_scan(run, L, part);
      for (uint32_t f = t; f < L; f += blockDim.x) gbase[f] = b0;
    }
    __syncthreads();
    for_each_entry(mid, s0, cnt, [&](uint2 e) { sorted[atomicAdd(&run[e.x & mask], 1u)] = e; });
    __syncthreads();
    for (uint32_t i = t; i < cnt; i += blockDim.x) {
      uint2 e = sorted[i];
      uint32_t p = gbase[e.x & mask] + i;
      keys_out[p] = e.x;
      vals_out[p] = e.y;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------- host side
static uint32_t pow2_ceil(uint64_t x) {
  uint32_t b = 0;
  while ((1ull << b) < x) b++;
  return b;
}

SortPlan msm_sort_plan(const MsmPlan& p) {
  SortPlan s{};
  const uint32_t kb = pow2_ceil(p.n_buckets);              // key space rounded up to a power of two
  // coarse bins: enough that small inputs get private bins (average fill <= a third of a segment),
  // never more than kMaxCoarse (run length of pass B), and enough that a bin has <= kMaxFine buckets
  uint32_t hb = pow2_ceil((p.m + kCap / 6 - 1) / (kCap / 6));
  const uint32_t hb_min = kb > 12 ? kb - 12 : 0, hb_max = 11;
  if (hb > hb_max) hb = hb_max;
  if (hb < hb_min) hb = hb_min;
  if (hb > kb) hb = kb;
  s.lo_bits = kb - hb;
  s.n_coarse = 1u << hb;
  s.max_items = s.n_coarse + uint32_t(p.m / (kCap / 2)) + 1;
  s.fine_entries = uint64_t(s.n_coarse) << s.lo_bits;
  return s;
}

size_t msm_sort_scratch_bytes(const MsmPlan& p) {
  SortPlan s = msm_sort_plan(p);
  // coarse_hist | coarse_off (+1) | cursor | bin_multi | n_items | items | fine
  return size_t(s.n_coarse) * 4 * 4 + 64 + size_t(s.max_items) * 8 + s.fine_entries * 4 + 256;
}

void msm_digits_and_bucket_sort(const MsmPlan& p, const MsmBuffers& b, const void* d_scalars, uint64_t n,
                                uint32_t idx_base, int sm_count, cudaStream_t st) {
  const SortPlan s = msm_sort_plan(p);
  const size_t place_smem = size_t(kSeg) * 8 + (size_t(8) << s.lo_bits);
  const size_t scatter_smem = size_t(kTile) * 8 + size_t(8) * s.n_coarse;
  // opt-in to large dynamic shared memory, once per device of this process
  static std::once_flag configured[64];
  int dev = 0;
  cudaGetDevice(&dev);
  std::call_once(configured[dev & 63], [] {
    cudaFuncSetAttribute(sort_place_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(size_t(kSeg) * 8 + kMaxFine * 8));
    const int scatter_max = int(size_t(kTile) * 8 + kMaxCoarse * 8);
    cudaFuncSetAttribute(sort_scatter_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, scatter_max);
    cudaFuncSetAttribute(sort_scatter_kernel<15>, cudaFuncAttributeMaxDynamicSharedMemorySize, scatter_max);
    cudaFuncSetAttribute(sort_scatter_kernel<17>, cudaFuncAttributeMaxDynamicSharedMemorySize, scatter_max);
    cudaFuncSetAttribute(sort_scatter_kernel<20>, cudaFuncAttributeMaxDynamicSharedMemorySize, scatter_max);
    cudaFuncSetAttribute(sort_scatter_kernel<22>, cudaFuncAttributeMaxDynamicSharedMemorySize, scatter_max);
    cudaFuncSetAttribute(sort_scatter_kernel<24>, cudaFuncAttributeMaxDynamicSharedMemorySize, scatter_max);
  });
  if (b.ev) cudaEventRecord(b.ev[0], st);
  char* sc = reinterpret_cast<char*>(b.sort_temp);
  uint32_t* coarse_hist = reinterpret_cast<uint32_t*>(sc);
  uint32_t* coarse_off = coarse_hist + s.n_coarse;
  uint32_t* cursor = coarse_off + s.n_coarse + 1;
  uint32_t* bin_multi = cursor + s.n_coarse;
  uint32_t* n_items = bin_multi + s.n_coarse;
  uint2* items = reinterpret_cast<uint2*>((reinterpret_cast<uintptr_t>(n_items + 4) + 15) & ~uintptr_t(15));
  uint32_t* fine = reinterpret_cast<uint32_t*>(items + s.max_items);
  uint2* mid = reinterpret_cast<uint2*>(b.keys_in);          // 8 bytes per entry (ensure_msm_buffers)
  cudaMemsetAsync(coarse_hist, 0, size_t(s.n_coarse) * 4, st);
  cudaMemsetAsync(fine, 0, s.fine_entries * 4, st);
  DigitCfg d;
  d.c = uint32_t(p.c);
  d.nwin = p.nwin;
  d.half = p.half;
  d.mask = (1u << p.c) - 1;
  d.val_stride = uint32_t(p.val_stride);
  d.idx_base = idx_base;
  d.one_set = p.val_stride != 0;
  // pass A: any tiling does; pass B: a tile's entries (<= windows per scalar) must fit the staging area
  const uint32_t per_a = 4096, per_b = kTile / p.nwin;
  const uint32_t tiles_a = uint32_t((n + per_a - 1) / per_a), tiles_b = uint32_t((n + per_b - 1) / per_b);
  const uint32_t* sc32 = reinterpret_cast<const uint32_t*>(d_scalars);
  // the widths msm_auto_window picks get unrolled digit loops; any other width takes the generic kernels
#define ZKB_SORT_AB(C)                                                                                              \
  do {                                                                                                              \
    sort_hist_kernel<C><<<tiles_a ? tiles_a : 1, kSortThreads, size_t(s.n_coarse) * 4, st>>>(                       \
        sc32, n, d, s.lo_bits, s.n_coarse, per_a, coarse_hist, b.flag);                                             \
    sort_plan_kernel<<<1, 1024, 0, st>>>(coarse_hist, s.n_coarse, coarse_off, cursor, b.m_valid, items, n_items,    \
                                         bin_multi);                                                                \
    sort_scatter_kernel<C><<<tiles_b ? tiles_b : 1, kPlaceThreads, scatter_smem, st>>>(                             \
        sc32, n, d, s.lo_bits, s.n_coarse, per_b, cursor, mid);                                                     \
  } while (0)
  switch (p.c) {
    case 15: ZKB_SORT_AB(15); break;
    case 17: ZKB_SORT_AB(17); break;
    case 20: ZKB_SORT_AB(20); break;
    case 22: ZKB_SORT_AB(22); break;
    case 24: ZKB_SORT_AB(24); break;
    default: ZKB_SORT_AB(0); break;
  }
#undef ZKB_SORT_AB
  if (b.ev) cudaEventRecord(b.ev[1], st);
  const uint32_t grid_c = uint32_t(sm_count) * 4 < s.max_items ? uint32_t(sm_count) * 4 : s.max_items;
  sort_fine_hist_kernel<<<grid_c, kSortThreads, 0, st>>>(mid, coarse_off, items, n_items, bin_multi, s.lo_bits, fine);
  sort_fine_scan_kernel<<<s.n_coarse < 296 ? s.n_coarse : 296, 1024, 0, st>>>(coarse_off, bin_multi, s.n_coarse, s.lo_bits, fine);
  const uint32_t grid_p = uint32_t(sm_count) * 2 < s.max_items ? uint32_t(sm_count) * 2 : s.max_items;
  sort_place_kernel<<<grid_p, kPlaceThreads, place_smem, st>>>(
      mid, coarse_off, items, n_items, bin_multi, s.lo_bits, fine, b.keys_out, b.vals_out);
  count_launches(6);
  if (b.ev) cudaEventRecord(b.ev[2], st);
}

}  // namespace zkb
