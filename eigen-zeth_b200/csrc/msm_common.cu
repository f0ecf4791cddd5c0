// msm_common.cu — the curve-independent front half of the MSM pipeline: signed-digit extraction,
// the (key, value) radix sort and the launch plan.  See msm_impl.cuh for the pipeline overview.
#include <cuda_runtime.h>

#include <cub/device/device_radix_sort.cuh>

#include "msm.h"

#include <atomic>

namespace zkb {

static std::atomic<uint64_t> g_launches{0};
void count_launches(uint64_t k) { g_launches.fetch_add(k, std::memory_order_relaxed); }
uint64_t launches_so_far() { return g_launches.load(std::memory_order_relaxed); }
static std::atomic<int> g_g2_pair{2};   // bit 0: lane-pair walk (5 % slower on B200), bit 1: lane-pair tails
static std::atomic<int> g_sort_mode{0};
void msm_set_sort_mode(int mode) { g_sort_mode.store(mode ? 1 : 0); }
int msm_get_sort_mode() { return g_sort_mode.load(); }
void msm_set_g2_pair(int on) { g_g2_pair.store(on & 3); }
int msm_get_g2_pair() { return g_g2_pair.load(); }

// ---------------------------------------------------------------- 1. digits
__global__ void msm_digits_kernel(const uint32_t* __restrict__ scalars, uint64_t n, uint32_t c,
                                  uint32_t nwin, uint32_t* __restrict__ keys,
                                  uint32_t* __restrict__ vals, uint32_t idx_base,
                                  uint32_t* __restrict__ flag, uint32_t val_stride) {
  uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const uint4* sp = reinterpret_cast<const uint4*>(scalars + 8 * i);
  uint4 lo = __ldg(sp), hi = __ldg(sp + 1);
  uint32_t s[9] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w, 0};
  {  // canonicity: s < r  (compare from the top limb)
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
    if (!lt) atomicOr(flag, 1u);
  }
  const uint32_t half = 1u << (c - 1);
  const uint32_t mask = (1u << c) - 1;
  const bool one_set = val_stride != 0;       // precomputed window multiples: a single bucket set
  const uint32_t sentinel = (one_set ? 1u : nwin) * half;
  uint32_t carry = 0;
  for (uint32_t w = 0; w < nwin; w++) {
    uint32_t bit = w * c;
    uint32_t word = bit >> 5, sh = bit & 31;
    uint32_t d = 0;
    if (word < 8) {
      uint64_t two = (uint64_t(s[word + 1]) << 32) | s[word];
      d = uint32_t(two >> sh) & mask;
    }
    d += carry;
    uint32_t neg = d > half;
    uint32_t mag = neg ? (1u << c) - d : d;
    carry = neg;
    uint32_t key = mag ? (one_set ? 0u : w * half) + (mag - 1) : sentinel;
    keys[uint64_t(w) * n + i] = key;
    vals[uint64_t(w) * n + i] = (w * val_stride + idx_base + uint32_t(i)) | (neg << 31);
  }
}

// first index whose key is >= limit in a sorted array (single thread, binary search)
__global__ void msm_count_valid_kernel(const uint32_t* keys, uint64_t m, uint32_t limit,
                                       uint64_t* out) {
  uint64_t lo = 0, hi = m;
  while (lo < hi) {
    uint64_t mid = (lo + hi) >> 1;
    if (keys[mid] < limit) lo = mid + 1; else hi = mid;
  }
  *out = lo;
}

size_t msm_sort_temp_bytes(uint64_t m, int end_bit) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                  (const uint32_t*)nullptr, (uint32_t*)nullptr, (int64_t)m, 0,
                                  end_bit);
  return bytes;
}

int msm_auto_window(uint64_t n, bool precomputed) {
  int log_n = 0;
  while ((1ull << log_n) < n) log_n++;
  if (precomputed) {
    // ONE bucket set shared by all windows: the width can grow until the bucket reduction
    // (2 adds per bucket) shows.  Only the smallest c of each window count ceil(255/c) is useful.
    // Calibrated on B200 (tools/tune_window.py, profiles/r01_tune_window.json; re-checked on whole proofs
    // with the batched / concurrent tails of round 2: tools/tune_prove_window.py).
    if (log_n <= 14) return 11;   // 24 windows
    if (log_n <= 16) return 15;   // 17
    if (log_n <= 20) return 17;   // 15  (2^20: 17 beats 20 by 4 % inside a proof, profiles/r02_tune_prove_window.json)
    if (log_n <= 22) return 20;   // 13
    return 22;                    // 12
  }
  // classic: one bucket set per window, keep buckets << points
  int c = log_n - 6;
  if (c < 4) c = 4;
  if (c > 16) c = 16;
  return c;
}

MsmPlan msm_make_plan(uint64_t n, int is_g2, int c_override, int sm_count, uint64_t precomp_stride) {
  MsmPlan p{};
  int c = c_override > 0 ? c_override : msm_auto_window(n, precomp_stride != 0);
  p.c = c;
  p.nwin = (255 + c - 1) / c;
  p.half = 1u << (c - 1);
  p.n_sets = precomp_stride ? 1 : p.nwin;
  p.val_stride = precomp_stride;
  p.n_buckets = uint64_t(p.n_sets) * p.half;
  p.m = n * p.nwin;
  p.key_bits = 1;
  while ((1ull << p.key_bits) <= p.n_buckets) p.key_bits++;   // sentinel == n_buckets must fit
  // level-1 threads: one resident wave, never less than 8 entries per thread
  p.g2_pair = is_g2 ? msm_get_g2_pair() : 0;
  p.sort_mode = msm_get_sort_mode();
  p.sm_count = sm_count;
  uint32_t per_sm = is_g2 ? msm_l1_threads_per_sm_g2(p.g2_pair) : msm_l1_threads_per_sm_g1();
  uint64_t t1 = uint64_t(sm_count) * per_sm;
  uint64_t cap = (p.m + 7) / 8;
  if (cap < 1) cap = 1;
  if (t1 > cap) t1 = cap;
  p.l1_threads = uint32_t(t1);
  p.l2_len = 32;
  p.l2_threads = (2 * p.l1_threads + p.l2_len - 1) / p.l2_len;
  // one 16-way level first only when the bucket array is huge (msm_impl.cuh, section 6)
  p.red_levels = p.half >= (1u << 20) ? 2 : 1;
  p.red_run_total = 0;
  for (uint32_t l = 0, cnt = p.half; l < p.red_levels; l++) {
    uint32_t groups = (cnt + 15) / 16;
    p.red_run_total += uint64_t(groups) * p.n_sets;
    cnt = groups;
  }
  p.point_bytes = is_g2 ? 256 : 128;
  p.sort_temp_bytes = msm_sort_temp_bytes(p.m, p.key_bits);
  const size_t own = msm_sort_scratch_bytes(p);
  if (own > p.sort_temp_bytes) p.sort_temp_bytes = own;
  return p;
}


uint64_t msm_tail_scratch_points(const MsmPlan& p) {
  uint64_t stage = 0;
  uint32_t cnt = p.half;
  for (uint32_t l = 0; l < p.red_levels; l++) {
    cnt = (cnt + 15) / 16;
    stage += uint64_t(msm_tail_slices(cnt)) * p.n_sets;       // this level's w0
  }
  stage += uint64_t(msm_tail_slices(cnt)) * p.n_sets;         // the last level's runs
  return 2 * p.red_run_total + stage;
}

void msm_digits_and_sort(const MsmPlan& p, const MsmBuffers& b, const void* d_scalars, uint64_t n,
                         uint32_t idx_base, cudaStream_t s) {
  if (p.sort_mode == 0) {
    msm_digits_and_bucket_sort(p, b, d_scalars, n, idx_base, p.sm_count, s);
    return;
  }
  if (b.ev) cudaEventRecord(b.ev[0], s);
  uint32_t blocks = uint32_t((n + 255) / 256);
  if (blocks)
    msm_digits_kernel<<<blocks, 256, 0, s>>>(reinterpret_cast<const uint32_t*>(d_scalars), n, p.c,
                                             p.nwin, b.keys_in, b.vals_in, idx_base, b.flag,
                                             uint32_t(p.val_stride));
  if (b.ev) cudaEventRecord(b.ev[1], s);
  size_t temp = p.sort_temp_bytes;
  cub::DeviceRadixSort::SortPairs(b.sort_temp, temp, (const uint32_t*)b.keys_in, b.keys_out,
                                  (const uint32_t*)b.vals_in, b.vals_out, (int64_t)p.m, 0,
                                  p.key_bits, s);
  msm_count_valid_kernel<<<1, 1, 0, s>>>(b.keys_out, p.m, uint32_t(p.n_buckets), b.m_valid);
  count_launches(blocks ? 2 : 1);
  if (b.ev) cudaEventRecord(b.ev[2], s);
}

}  // namespace zkb
