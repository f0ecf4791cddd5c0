// msm.h — plan / scratch descriptors shared by msm.cu and the context.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace zkb {

// kernels of this library launched by the process so far (every launch site adds its own count;
// library kernels such as the CUB sort are not included)
void count_launches(uint64_t k);
uint64_t launches_so_far();

struct MsmPlan {
  int c;                    // window width (signed digits in [-2^{c-1}, 2^{c-1}])
  uint32_t nwin;            // ceil(255 / c)
  uint32_t n_sets;          // bucket sets: nwin, or 1 when the bases carry precomputed window multiples
  uint64_t val_stride;      // precomputed bases: window w reads point w * val_stride + i (0 otherwise)
  uint32_t half;            // buckets per window = 2^{c-1}
  uint64_t n_buckets;       // n_sets * half (also the sentinel key of zero digits)
  uint64_t m;               // n * nwin (key, value) pairs
  int key_bits;             // radix-sort key width
  int g2_pair;              // G2 only: bit 0 lane-pair walk, bit 1 lane-pair tail kernels
  int sort_mode;            // 0 = hand-written bucket sort, 1 = cub radix sort
  int sm_count;
  uint32_t l1_threads;      // chunks of the level-1 walk
  uint32_t l2_len;          // chunk length of the level-2 walk
  uint32_t l2_threads;
  uint32_t red_levels;      // bucket-reduction levels (1, or 2 for >= 2^20 buckets per set)
  uint64_t red_run_total;   // points of run output over all levels and sets
  size_t point_bytes;       // sizeof(XYZZ)
  size_t sort_temp_bytes;
};

struct SortPlan {            // geometry of the hand-written bucket sort (msm_sort.cu)
  uint32_t lo_bits;          // fine bucket bits inside a coarse bin
  uint32_t n_coarse;         // coarse bins
  uint32_t max_items;        // upper bound of pass-C work items
  uint64_t fine_entries;     // n_coarse << lo_bits
};

struct MsmBuffers {
  uint32_t *keys_in, *vals_in, *keys_out, *vals_out;  // m entries each
  void* sort_temp;
  uint64_t* m_valid;
  void* buckets;            // n_buckets XYZZ
  void* buckets2;           // second bucket set (chunked MSM), may be null
  uint32_t* flag;           // bit 0: a scalar was not canonical
  uint32_t* part1_keys;     // 2 * l1_threads
  void* part1_pts;
  uint32_t* part2_keys;     // 2 * l2_threads
  void* part2_pts;
  void* group_sums;         // reduction scratch (see msm_tail_t)
  void* window_sums;        // (red_levels + 2) * n_sets XYZZ
  void* result_xyzz;        // 1 XYZZ (Montgomery)
  void* result_affine;      // 1 affine, canonical little-endian
  cudaEvent_t* ev;          // optional 6 profiling events: start, digits, sort, L1, L2+finish, tail
};

// precomp_stride != 0: the bases hold 2^{c w} P_i at index w * precomp_stride + i for every window w
// (c_override is then the width they were built for) and all windows share one bucket set.
MsmPlan msm_make_plan(uint64_t n, int is_g2, int c_override, int sm_count, uint64_t precomp_stride = 0);
int msm_auto_window(uint64_t n, bool precomputed);
// tail (msm_tail_t): partial sums a point array of `count` entries per set is cut into, and the XYZZ
// points of b.group_sums the plan's tail needs (runs + w0 of every level + the partials)
inline uint32_t msm_tail_slices(uint32_t count) {
  return count >= 65536 ? 256 : (count >= 4096 ? 64 : (count >= 256 ? 8 : 1));
}
uint64_t msm_tail_scratch_points(const MsmPlan& p);
// resident threads per SM the level-1 walk kernel of each curve is compiled for (launch bounds)
uint32_t msm_l1_threads_per_sm_g1();
uint32_t msm_l1_threads_per_sm_g2(int pair);
// process-wide comparison switch (zkb_ctx_set "msm_g2_pair"): bit 0 = lane-pair G2 walk, bit 1 = lane-pair
// latency-bound G2 kernels (default 2)
void msm_set_g2_pair(int on);
int msm_get_g2_pair();
// out[w * n + i] = 2^{c w} * in[i], w < nwin (affine Montgomery in and out; in may alias out[0..n))
void msm_precompute_g1(const void* d_in, void* d_out, uint64_t n, uint32_t nwin, uint32_t c, cudaStream_t s);
void msm_precompute_g2(const void* d_in, void* d_out, uint64_t n, uint32_t nwin, uint32_t c, cudaStream_t s);
// scalars[0..n) -> sorted (key, value) stream; values carry idx_base + i.  Non-canonical scalars
// (>= r) raise bit 0 of *d_flag.
void msm_digits_and_sort(const MsmPlan& p, const MsmBuffers& b, const void* d_scalars, uint64_t n,
                         uint32_t idx_base, cudaStream_t s);
// the same contract, hand-written: digit extraction fused with a two-level MSD bucket sort (msm_sort.cu);
// needs b.keys_in of 8 bytes per entry and b.sort_temp of msm_sort_scratch_bytes(p)
void msm_digits_and_bucket_sort(const MsmPlan& p, const MsmBuffers& b, const void* d_scalars, uint64_t n,
                                uint32_t idx_base, int sm_count, cudaStream_t s);
SortPlan msm_sort_plan(const MsmPlan& p);
size_t msm_sort_scratch_bytes(const MsmPlan& p);
// process-wide switch (zkb_ctx_set "msm_sort"): 0 = the bucket sort above (default), 1 = digits kernel +
// cub::DeviceRadixSort (library baseline, kept for comparison)
void msm_set_sort_mode(int mode);
int msm_get_sort_mode();
void msm_accumulate_g1(const MsmPlan& p, const MsmBuffers& b, const void* d_bases, void* bucket_target,
                       cudaStream_t s);
void msm_accumulate_g2(const MsmPlan& p, const MsmBuffers& b, const void* d_bases, void* bucket_target,
                       cudaStream_t s);
// the two halves of msm_accumulate_*: the level-1 walk (throughput-bound, owns the GPU) and the
// reduction of its partial list into the buckets (short latency-bound launches that can share the
// GPU with those of other MSMs on other streams); both use b.part1_* / b.part2_*
void msm_walk_g1(const MsmPlan& p, const MsmBuffers& b, const void* d_bases, void* bucket_target, cudaStream_t s);
void msm_walk_g2(const MsmPlan& p, const MsmBuffers& b, const void* d_bases, void* bucket_target, cudaStream_t s);
void msm_post_g1(const MsmPlan& p, const MsmBuffers& b, void* bucket_target, cudaStream_t s);
void msm_post_g2(const MsmPlan& p, const MsmBuffers& b, void* bucket_target, cudaStream_t s);
void msm_merge_g1(const MsmPlan& p, void* main_buckets, const void* part_buckets, cudaStream_t s);
void msm_merge_g2(const MsmPlan& p, void* main_buckets, const void* part_buckets, cudaStream_t s);
void msm_tail_g1(const MsmPlan& p, const MsmBuffers& b, cudaStream_t s);
void msm_tail_g2(const MsmPlan& p, const MsmBuffers& b, cudaStream_t s);
// ONE tail over p.n_sets (<= 8) bucket sets that belong to different MSMs (b.buckets = [set][half]);
// set i writes its canonical affine result to d_outs[i]
void msm_tail_multi_g1(const MsmPlan& p, const MsmBuffers& b, void* const* d_outs, uint32_t n, cudaStream_t s);
// canonical affine -> Montgomery affine in place; every point that is neither the all-zero encoding
// of infinity nor on the curve raises bit 1 of *d_flag (d_flag may be null: no check).
// b_mont: the curve coefficient (8 / 16 Montgomery words, host pointer).
void msm_bases_to_mont_g1(void* d_pts, uint64_t n, const uint32_t* b_mont, uint32_t* d_flag, cudaStream_t s);
void msm_bases_to_mont_g2(void* d_pts, uint64_t n, const uint32_t* b_mont, uint32_t* d_flag, cudaStream_t s);
void host_curve_b(int is_g2, uint32_t* out);
void msm_sum_affine_g1(const void* d_pts, uint32_t n, void* d_out, cudaStream_t s);
void msm_sum_affine_g2(const void* d_pts, uint32_t n, void* d_out, cudaStream_t s);

}  // namespace zkb
