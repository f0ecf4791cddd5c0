// msm_impl.cuh — signed-window Pippenger multi-scalar multiplication for BN254 G1 and G2 on sm_100a.
//
// No reference counterpart exists: these are the A/B1/L/H G1 MSMs
// and the B2 G2 MSM of the Groth16 step behind reference proto/prover/v1/prover.proto:130-148.
//
// Pipeline (all on device, one stream):
//   1. msm_digits        scalars -> signed radix-2^c digits; key = window*2^{c-1} + |d|-1,
//                        value = point index | sign<<31; zero digits get the sentinel key.
//   2. radix sort        (key,value) pairs by key (cub::DeviceRadixSort, key bits only).
//   3. accumulate L1     the sorted stream is cut into equal chunks, one per thread; a thread
//                        walks its chunk adding gathered affine points into an XYZZ accumulator.
//                        Runs that lie strictly inside a chunk are whole buckets and are written
//                        straight to the bucket array; the first and last run of a chunk are
//                        emitted as (key, partial sum) pairs.  Work per thread is identical
//                        whatever the bucket-size distribution is.
//   4. accumulate L2     the same chunked walk over the partial list (XYZZ + XYZZ).
//   5. finish            one thread per remaining run.
//   6. bucket reduce     levels of 16-way running sums (no scalar multiplications, see below).
//   7. window sum        tree reduction of each level's weighted sums.
//   8. Horner            over the reduce levels (x16) and the bucket sets (x2^c), to affine.
#pragma once
#include <cuda_runtime.h>

#include <cstring>

#include "ec.cuh"
#include "msm.h"

namespace zkb {

constexpr uint32_t kInvalidKey = 0xffffffffu;

// resident CTAs (of 128 threads) per SM for the level-1 walk: 4 x 128 threads at <= 128 registers
// for G1 and for G2 split over lane pairs (Fq2L, ff.cuh); the one-thread-per-point Fq2 walk needs the
// full 255 registers, so 2 x 128 (kept for comparison: ctx tunable "msm_g2_pair" = 0)
template <class F> struct L1Cfg { static constexpr int kMinBlocks = 4; };
template <> struct L1Cfg<Fq2> { static constexpr int kMinBlocks = 2; };

// ---------------------------------------------------------------- 3. accumulate, level 1
// F is the register-level field type of the walk (Fq, Fq2, or Fq2L = one Fq2 over a lane pair);
// memory is addressed in bytes so that Fq2 and Fq2L share one layout.  A chunk is walked by
// F::kLanes adjacent lanes.
template <class F>
__global__ void __launch_bounds__(128, L1Cfg<F>::kMinBlocks) msm_accumulate_l1_kernel(
    const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals,
    const uint64_t* __restrict__ m_valid_ptr, const char* __restrict__ bases,
    char* __restrict__ buckets, uint32_t* __restrict__ part_keys,
    char* __restrict__ part_pts, uint32_t n_threads) {
  constexpr size_t kAff = 2 * size_t(F::kBytes), kPt = 4 * size_t(F::kBytes);
  uint32_t tid = (blockIdx.x * blockDim.x + threadIdx.x) / F::kLanes;
  if (tid >= n_threads) return;
  const uint64_t m_valid = *m_valid_ptr;
  const uint64_t len = (m_valid + n_threads - 1) / n_threads;
  uint64_t start = uint64_t(tid) * len;
  uint64_t end = start + len < m_valid ? start + len : m_valid;
  if (start >= end) {
    part_keys[2 * tid] = kInvalidKey;
    part_keys[2 * tid + 1] = kInvalidKey;
    return;
  }
  uint32_t key_n = keys[start], val_n = vals[start];
  Affine<F> pt_n = Affine<F>::load_nc(bases + size_t(val_n & 0x7fffffffu) * kAff);
  uint32_t cur = key_n;
  bool first_run = true;
  XYZZ<F> acc = XYZZ<F>::infinity();
  for (uint64_t i = start; i < end; i++) {
    uint32_t key = key_n, val = val_n;
    Affine<F> pt = pt_n;
    if (i + 1 < end) {
      key_n = keys[i + 1];
      val_n = vals[i + 1];
      pt_n = Affine<F>::load_nc(bases + size_t(val_n & 0x7fffffffu) * kAff);
    }
    if (key != cur) {
      if (first_run) {
        part_keys[2 * tid] = cur;
        acc.store(part_pts + size_t(2 * tid) * kPt);
        first_run = false;
      } else {
        acc.store(buckets + size_t(cur) * kPt);
      }
      acc = XYZZ<F>::infinity();
      cur = key;
    }
    if (val >> 31) pt.y = pt.y.neg();
    acc.madd(pt);
  }
  if (first_run) {
    part_keys[2 * tid] = cur;
    acc.store(part_pts + size_t(2 * tid) * kPt);
    part_keys[2 * tid + 1] = kInvalidKey;
  } else {
    part_keys[2 * tid + 1] = cur;
    acc.store(part_pts + size_t(2 * tid + 1) * kPt);
  }
}

// The kernels below (partial merges, bucket reduction, window sums, Horner) are short chains of
// dependent point additions on small grids: latency-bound.  They address memory in bytes and let
// F::kLanes adjacent lanes share one work item, so G2 can run them on Fq2L (one Fq2 operation over a
// lane pair: half the dependent IMAD.WIDE chain per lane, twice the warps on otherwise idle SMs).
template <class F> __device__ __forceinline__ XYZZ<F> ld_pt(const char* base, uint64_t idx) {
  return XYZZ<F>::load(base + idx * (4 * size_t(F::kBytes)));
}
template <class F> __device__ __forceinline__ void st_pt(const XYZZ<F>& v, char* base, uint64_t idx) {
  v.store(base + idx * (4 * size_t(F::kBytes)));
}

// ---------------------------------------------------------------- 4. accumulate, level 2
// Same walk over a (key, XYZZ) list that may contain kInvalidKey holes.
template <class F>
__global__ void __launch_bounds__(128) msm_accumulate_l2_kernel(
    const uint32_t* __restrict__ keys, const char* __restrict__ pts, uint32_t m, uint32_t len,
    char* __restrict__ buckets, uint32_t* __restrict__ part_keys,
    char* __restrict__ part_pts, uint32_t n_threads) {
  uint32_t tid = (blockIdx.x * blockDim.x + threadIdx.x) / F::kLanes;
  if (tid >= n_threads) return;
  uint32_t start = tid * len;
  uint32_t end = start + len < m ? start + len : m;
  uint32_t cur = kInvalidKey;
  bool first_run = true;
  XYZZ<F> acc = XYZZ<F>::infinity();
  part_keys[2 * tid] = kInvalidKey;
  part_keys[2 * tid + 1] = kInvalidKey;
  for (uint32_t i = start; i < end; i++) {
    uint32_t key = keys[i];
    if (key == kInvalidKey) continue;
    if (cur == kInvalidKey) cur = key;
    if (key != cur) {
      if (first_run) {
        part_keys[2 * tid] = cur;
        st_pt<F>(acc, part_pts, 2 * tid);
        first_run = false;
      } else {
        st_pt<F>(acc, buckets, cur);
      }
      acc = XYZZ<F>::infinity();
      cur = key;
    }
    acc.add(ld_pt<F>(pts, i));
  }
  if (cur == kInvalidKey) return;
  if (first_run) {
    part_keys[2 * tid] = cur;
    st_pt<F>(acc, part_pts, 2 * tid);
  } else {
    part_keys[2 * tid + 1] = cur;
    st_pt<F>(acc, part_pts, 2 * tid + 1);
  }
}

// ---------------------------------------------------------------- 5. finish
// One work item per list entry; the item that holds the first valid entry of a key sums the run.
template <class F>
__global__ void __launch_bounds__(128) msm_finish_kernel(const uint32_t* __restrict__ keys,
                                                         const char* __restrict__ pts,
                                                         uint32_t m, char* __restrict__ buckets) {
  uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) / F::kLanes;
  if (i >= m) return;
  uint32_t key = keys[i];
  if (key == kInvalidKey) return;
  for (uint32_t j = i; j > 0; j--) {     // previous valid entry with the same key -> not the head
    uint32_t kp = keys[j - 1];
    if (kp == kInvalidKey) continue;
    if (kp == key) return;
    break;
  }
  XYZZ<F> acc = ld_pt<F>(pts, i);
  for (uint32_t j = i + 1; j < m; j++) {
    uint32_t kn = keys[j];
    if (kn == kInvalidKey) continue;
    if (kn != key) break;
    acc.add(ld_pt<F>(pts, j));
  }
  st_pt<F>(acc, buckets, key);
}

// ---------------------------------------------------------------- 6. bucket reduce
// sum_j (j+1) B_j without any scalar multiplication: with zero-based weights W0(B) = sum_j j B_j
// and groups of 16,  W0(B) = sum_g w0_g + 16 * W0(run),  w0_g = sum_t t B_{16g+t},
// run_g = sum_t B_{16g+t}.  One level = one launch: item g turns 16 inputs into (run_g, w0_g)
// with 31 additions; the runs are the next level's input (16x fewer), the w0's of a level are
// summed by the window-sum kernel, and the answer is  S + sum_l 16^l W_l  (S = sum of the runs of
// the last level).  Each level is a latency-bound launch, so only arrays of >= 2^20 buckets take a
// 16-way level first; the (last) level then folds 16 g run_g in with a small scalar multiplication.
constexpr uint32_t kRedGroup = 16;

// resident CTAs per SM the reduce level is compiled for: a level over 2^19..2^21 buckets is
// throughput-bound, and unconstrained ptxas gives the G1 form 202 registers = 8 warps / SM; capped at
// 168 registers (a few spills) it runs 12 warps / SM.  The G2 forms need every register they can get.
template <class F> struct ReduceCfg { static constexpr int kMinBlocks = 1; };
template <> struct ReduceCfg<Fq> { static constexpr int kMinBlocks = 3; };

template <class F>
__global__ void __launch_bounds__(128, ReduceCfg<F>::kMinBlocks) msm_reduce_level_kernel(const char* __restrict__ in,
                                                               uint32_t count, uint32_t groups,
                                                               uint32_t n_sets, char* __restrict__ run_out,
                                                               char* __restrict__ w0_out, int last_level) {
  uint32_t gid = (blockIdx.x * blockDim.x + threadIdx.x) / F::kLanes;
  if (gid >= groups * n_sets) return;
  uint32_t set = gid / groups, g = gid - set * groups;
  const uint64_t b = uint64_t(set) * count + uint64_t(g) * kRedGroup;
  uint32_t avail = count - g * kRedGroup;
  if (avail > kRedGroup) avail = kRedGroup;
  XYZZ<F> r = XYZZ<F>::infinity(), w = XYZZ<F>::infinity();
  for (int t = int(avail) - 1; t >= 1; t--) {
    r.add(ld_pt<F>(in, b + t));
    w.add(r);
  }
  r.add(ld_pt<F>(in, b));
  // the last level closes the recursion with one small scalar multiplication per group:
  // sum_j j x_j over the whole array = sum_g (w0_g + 16 g run_g)
  if (last_level && g) w.add(mul_small(r, g * kRedGroup));
  st_pt<F>(r, run_out, gid);
  st_pt<F>(w, w0_out, gid);
}

// ---------------------------------------------------------------- 7. window sum
// One launch sums SEVERAL point arrays at once (the w0's of every reduce level and the runs of the
// last one: independent, latency-bound sums that used to be one launch pair each).  CTA (job, set,
// slice) sums its slice of the `count` points of a set (serial per item, then a shared memory tree
// over the blockDim.x / kLanes items) and writes point [set * slices + slice] of the job's output.
// Launched twice: inputs -> `slices` partials per set -> one point per set.
constexpr uint32_t kMaxSumJobs = 4;
struct SumJobs {
  const char* in[kMaxSumJobs];
  char* out[kMaxSumJobs];
  uint32_t count[kMaxSumJobs], slices[kMaxSumJobs], first[kMaxSumJobs + 1];   // first[j] = first CTA of job j
  uint32_t n;
};

template <class F>
__global__ void __launch_bounds__(128) msm_window_sum_kernel(SumJobs jobs) {
  extern __shared__ uint4 sm_raw[];
  char* sm = reinterpret_cast<char*>(sm_raw);
  uint32_t j = 0;
  while (j + 1 < jobs.n && blockIdx.x >= jobs.first[j + 1]) j++;
  const char* in = jobs.in[j];
  const uint32_t count = jobs.count[j], slices = jobs.slices[j], local = blockIdx.x - jobs.first[j];
  const uint32_t items = blockDim.x / F::kLanes;
  const uint32_t set = local / slices, slice = local - set * slices, t = threadIdx.x / F::kLanes;
  const uint32_t len = (count + slices - 1) / slices;
  const uint32_t lo = slice * len, hi = lo + len < count ? lo + len : count;
  XYZZ<F> acc = XYZZ<F>::infinity();
  for (uint32_t g = lo + t; g < hi; g += items) acc.add(ld_pt<F>(in, uint64_t(set) * count + g));
  st_pt<F>(acc, sm, t);
  __syncthreads();
  for (uint32_t s = items >> 1; s > 0; s >>= 1) {
    if (t < s) {
      XYZZ<F> a = ld_pt<F>(sm, t);
      a.add(ld_pt<F>(sm, t + s));
      st_pt<F>(a, sm, t);
    }
    __syncthreads();
  }
  if (t == 0) st_pt<F>(ld_pt<F>(sm, 0), jobs.out[j], local);
}

// ---------------------------------------------------------------- 8. Horner + output
// per set: S + sum_l 16^l W_l (Horner over the reduce levels, 4 doublings each), then Horner over
// the sets with c doublings each.  wl is [level][set]; s_tot is [set].
// result (Montgomery XYZZ, 4*F::kBytes bytes) and canonical affine (2*F::kBytes bytes)
// Independent mode (outs.n != 0): the sets are the bucket sets of DIFFERENT MSMs (one each, bases with
// precomputed window multiples) reduced by one shared tail; set t writes its own canonical affine
// result to outs.p[t] and nothing is combined across sets.
struct MsmOutPtrs {
  void* p[8];
  uint32_t n;
};

template <class F>
__global__ void msm_horner_kernel(const char* __restrict__ wl, const char* __restrict__ s_tot,
                                  uint32_t levels, uint32_t n_sets, uint32_t c,
                                  char* __restrict__ set_sums, char* __restrict__ out_xyzz,
                                  char* __restrict__ out_affine, MsmOutPtrs outs) {
  const uint32_t t = threadIdx.x / F::kLanes;
  if (blockIdx.x) return;
  if (t < n_sets) {
    XYZZ<F> acc = ld_pt<F>(wl, uint64_t(levels - 1) * n_sets + t);
    for (int l = int(levels) - 2; l >= 0; l--) {
      for (int k = 0; k < 4; k++) acc = acc.dbl();
      acc.add(ld_pt<F>(wl, uint64_t(l) * n_sets + t));
    }
    acc.add(ld_pt<F>(s_tot, t));
    if (outs.n) {
      Affine<F> a = acc.to_affine();
      a.x = a.x.from_mont();
      a.y = a.y.from_mont();
      a.store(outs.p[t]);
    } else {
      st_pt<F>(acc, set_sums, t);
    }
  }
  if (outs.n) return;
  __syncthreads();
  if (t) return;
  XYZZ<F> acc = ld_pt<F>(set_sums, n_sets - 1);
  for (int w = int(n_sets) - 2; w >= 0; w--) {
    for (uint32_t i = 0; i < c; i++) acc = acc.dbl();
    acc.add(ld_pt<F>(set_sums, w));
  }
  acc.store(out_xyzz);
  Affine<F> a = acc.to_affine();
  a.x = a.x.from_mont();
  a.y = a.y.from_mont();
  a.store(out_affine);
}

// ---------------------------------------------------------------- base upload / point utilities
// canonical affine -> Montgomery affine (infinity = all zero stays all zero); points that are not on
// y^2 = x^3 + b raise bit 1 of *flag (an off-curve point of small even order would break the
// "a base never doubles to infinity" assumption of msm_precompute_kernel and silently zero its
// window multiples)
template <class F>
__global__ void msm_bases_to_mont_kernel(Affine<F>* pts, uint64_t n, F b, uint32_t* flag) {
  uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  Affine<F> p = Affine<F>::load(pts + i);
  const bool inf = p.is_inf();
  p.x = p.x.to_mont();
  p.y = p.y.to_mont();
  if (flag && !inf && !(p.y.sqr() == p.x.sqr() * p.x + b)) atomicOr(flag, 2u);
  p.store(pts + i);
}

// sum of n canonical affine points (tiny n: partial results of the ranks) -> canonical affine
template <class F>
__global__ void msm_sum_affine_kernel(const Affine<F>* pts, uint32_t n, Affine<F>* out) {
  if (threadIdx.x || blockIdx.x) return;
  XYZZ<F> acc = XYZZ<F>::infinity();
  for (uint32_t i = 0; i < n; i++) {
    Affine<F> p = Affine<F>::load(pts + i);
    p.x = p.x.to_mont();
    p.y = p.y.to_mont();
    acc.madd(p);
  }
  Affine<F> a = acc.to_affine();
  a.x = a.x.from_mont();
  a.y = a.y.from_mont();
  a.store(out);
}

// main[j] += part[j] for every bucket (merging the bucket sets of point chunks)
template <class F>
__global__ void __launch_bounds__(128) msm_bucket_merge_kernel(XYZZ<F>* __restrict__ main_b,
                                                               const XYZZ<F>* __restrict__ part,
                                                               uint64_t n_buckets) {
  uint64_t j = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
  if (j >= n_buckets) return;
  XYZZ<F> a = XYZZ<F>::load(part + j);
  if (a.is_inf()) return;
  XYZZ<F> m = XYZZ<F>::load(main_b + j);
  m.add(a);
  m.store(main_b + j);
}

// sorted (key, value) stream -> bucket sums written to `bucket_target` (cleared first).
// FW is the register-level field type of the level-1 walk, FT that of the latency-bound kernels
// after it (F itself, or a type with the same storage layout: Fq2C / Fq2L for F = Fq2).
// level-1 walk only: sorted stream -> whole buckets in `bucket_target` + 2 partials per chunk in b.part1_*
template <class F, class FW = F>
void msm_walk_t(const MsmPlan& p, const MsmBuffers& b, const void* d_bases, void* bucket_target, cudaStream_t s) {
  static_assert(F::kBytes == FW::kBytes, "walk types must share the storage layout");
  constexpr size_t kPt = 4 * size_t(F::kBytes);
  cudaMemsetAsync(bucket_target, 0, p.n_buckets * kPt, s);
  msm_accumulate_l1_kernel<FW><<<(p.l1_threads * FW::kLanes + 127) / 128, 128, 0, s>>>(
      b.keys_out, b.vals_out, b.m_valid, reinterpret_cast<const char*>(d_bases),
      reinterpret_cast<char*>(bucket_target), b.part1_keys, reinterpret_cast<char*>(b.part1_pts),
      p.l1_threads);
  count_launches(1);
  if (b.ev) cudaEventRecord(b.ev[3], s);
}

// the partial list of a walk -> buckets.  The list shrinks 16x per level-2 pass (chunks of l2_len
// entries -> 2 partials); passes repeat, ping-ponging between the two partial buffers, until the
// list is short enough for the serial-per-run finish.  A single giant bucket (0/1 witnesses,
// all-equal scalars) therefore costs a few short passes, never a long serial chain.
template <class F, class FT = F>
void msm_post_t(const MsmPlan& p, const MsmBuffers& b, void* bucket_target, cudaStream_t s) {
  static_assert(F::kBytes == FT::kBytes, "types must share the storage layout");
  constexpr uint32_t L = FT::kLanes;
  uint32_t* keys_a = b.part1_keys;
  uint32_t* keys_b = b.part2_keys;
  char* pts_a = reinterpret_cast<char*>(b.part1_pts);
  char* pts_b = reinterpret_cast<char*>(b.part2_pts);
  uint32_t count = 2 * p.l1_threads;
  do {
    uint32_t threads = (count + p.l2_len - 1) / p.l2_len;
    msm_accumulate_l2_kernel<FT><<<(threads * L + 127) / 128, 128, 0, s>>>(
        keys_a, pts_a, count, p.l2_len, reinterpret_cast<char*>(bucket_target), keys_b, pts_b, threads);
    count_launches(1);
    count = 2 * threads;
    uint32_t* tk = keys_a; keys_a = keys_b; keys_b = tk;
    char* tp = pts_a; pts_a = pts_b; pts_b = tp;
  } while (count > 1024);
  msm_finish_kernel<FT><<<(count * L + 127) / 128, 128, 0, s>>>(keys_a, pts_a, count,
                                                               reinterpret_cast<char*>(bucket_target));
  count_launches(1);
  if (b.ev) cudaEventRecord(b.ev[4], s);
}

// sorted (key, value) stream -> bucket sums written to `bucket_target` (cleared first).
// FW is the register-level field type of the level-1 walk, FT that of the latency-bound kernels
// after it (F itself, or a type with the same storage layout: Fq2C / Fq2L for F = Fq2).
template <class F, class FW = F, class FT = F>
void msm_accumulate_t(const MsmPlan& p, const MsmBuffers& b, const void* d_bases,
                      void* bucket_target, cudaStream_t s) {
  msm_walk_t<F, FW>(p, b, d_bases, bucket_target, s);
  msm_post_t<F, FT>(p, b, bucket_target, s);
}

template <class F>
void msm_merge_t(const MsmPlan& p, void* main_buckets, const void* part_buckets, cudaStream_t s) {
  using P = XYZZ<F>;
  msm_bucket_merge_kernel<F><<<uint32_t((p.n_buckets + 127) / 128), 128, 0, s>>>(
      reinterpret_cast<P*>(main_buckets), reinterpret_cast<const P*>(part_buckets), p.n_buckets);
  count_launches(1);
}

// bucket sums -> one point (Montgomery XYZZ + canonical affine)
// scratch layout in b.group_sums (points): [runs of all levels][w0 of all levels][stage: slices * n_sets
// per summed array] (msm_tail_scratch_points)
// b.window_sums: [levels][n_sets] W_l, [n_sets] S, [n_sets] set sums
// F: register-level type of the latency-bound launches; FB: of a reduce level with so many groups
// (>= kBigLevel work items) that it is throughput-bound (G2: lane pairs lose there, 2.87 vs 2.36 ms
// at 2^19 buckets, and win below, 1.18 vs 1.73 ms at 2^16)
constexpr uint32_t kBigLevel = 16384;
template <class F, class FB = F>
void msm_tail_t(const MsmPlan& p, const MsmBuffers& b, cudaStream_t s, const MsmOutPtrs* independent = nullptr) {
  static_assert(F::kBytes == FB::kBytes, "tail types must share the storage layout");
  constexpr size_t kPt = 4 * size_t(F::kBytes);
  constexpr uint32_t L = F::kLanes;
  char* scratch = reinterpret_cast<char*>(b.group_sums);
  char* wl = reinterpret_cast<char*>(b.window_sums);
  const char* in = reinterpret_cast<const char*>(b.buckets);
  uint32_t count = p.half;
  char* run_cursor = scratch;
  char* w0 = scratch + p.red_run_total * kPt;
  char* stage = w0 + p.red_run_total * kPt;
  const size_t smem = (128 / L) * kPt;
  // the sums of every level's w0 (-> W_l) and of the last level's runs (-> S) wait until all levels
  // ran, then share one launch per stage
  SumJobs first{}, second{};
  auto add_sum = [&](const char* src, uint32_t n_in, char* dst) {
    const uint32_t slices = msm_tail_slices(n_in);
    const uint32_t j = first.n++;
    first.in[j] = src;
    first.count[j] = n_in;
    first.slices[j] = slices;
    first.first[j + 1] = first.first[j] + p.n_sets * slices;
    first.out[j] = slices > 1 ? stage : dst;
    if (slices > 1) {
      const uint32_t k = second.n++;
      second.in[k] = stage;
      second.count[k] = slices;
      second.slices[k] = 1;
      second.first[k + 1] = second.first[k] + p.n_sets;
      second.out[k] = dst;
      stage += uint64_t(p.n_sets) * slices * kPt;
    }
  };
  uint32_t groups = 0;
  for (uint32_t l = 0; l < p.red_levels; l++) {
    groups = (count + kRedGroup - 1) / kRedGroup;
    uint32_t threads = groups * p.n_sets;
    const int last = l + 1 == p.red_levels;
    if (threads >= kBigLevel)
      msm_reduce_level_kernel<FB><<<(threads * FB::kLanes + 127) / 128, 128, 0, s>>>(in, count, groups, p.n_sets, run_cursor, w0, last);
    else
      msm_reduce_level_kernel<F><<<(threads * L + 127) / 128, 128, 0, s>>>(in, count, groups, p.n_sets, run_cursor, w0, last);
    add_sum(w0, groups, wl + uint64_t(l) * p.n_sets * kPt);          // W_l[set] = sum of this level's w0
    in = run_cursor;
    run_cursor += uint64_t(threads) * kPt;
    w0 += uint64_t(threads) * kPt;
    count = groups;
  }
  add_sum(in, groups, wl + uint64_t(p.red_levels) * p.n_sets * kPt);  // S[set] = sum of the runs of the last level
  char* s_tot = wl + uint64_t(p.red_levels) * p.n_sets * kPt;
  msm_window_sum_kernel<F><<<first.first[first.n], 128, smem, s>>>(first);
  if (second.n) msm_window_sum_kernel<F><<<second.first[second.n], 128, smem, s>>>(second);
  count_launches(p.red_levels + (second.n ? 3 : 2));   // reduce levels + sums + the Horner kernel below
  MsmOutPtrs outs{};
  if (independent) outs = *independent;
  msm_horner_kernel<F><<<1, 128 * L, 0, s>>>(wl, s_tot, p.red_levels, p.n_sets, p.c,
                                         wl + uint64_t(p.red_levels + 1) * p.n_sets * kPt,
                                         reinterpret_cast<char*>(b.result_xyzz),
                                         reinterpret_cast<char*>(b.result_affine), outs);
  if (b.ev) cudaEventRecord(b.ev[5], s);
}

// ---------------------------------------------------------------- precomputed window multiples
// out[w * n + i] = 2^{c w} P_i.  One thread per base: a chain of c doublings per window in XYZZ,
// all windows brought back to affine with ONE inversion (Montgomery's trick over the chain).
constexpr int kMaxPrecompWindows = 32;

template <class F>
__global__ void __launch_bounds__(128) msm_precompute_kernel(const Affine<F>* __restrict__ in,
                                                             Affine<F>* __restrict__ out, uint64_t n,
                                                             uint32_t nwin, uint32_t c) {
  uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  Affine<F> p0 = Affine<F>::load(in + i);
  p0.store(out + i);
  if (p0.is_inf()) {
    for (uint32_t w = 1; w < nwin; w++) Affine<F>{F::zero(), F::zero()}.store(out + uint64_t(w) * n + i);
    return;
  }
  // a point of odd prime order never doubles to infinity, so every zzz below is invertible
  XYZZ<F> pts[kMaxPrecompWindows];
  F prefix[kMaxPrecompWindows];
  XYZZ<F> acc = XYZZ<F>::from_affine(p0);
  F run = F::one();
  for (uint32_t w = 1; w < nwin; w++) {
    for (uint32_t k = 0; k < c; k++) acc = acc.dbl();
    pts[w] = acc;
    prefix[w] = run;              // product of zzz_1 .. zzz_{w-1}
    run = run * acc.zzz;
  }
  F inv = run.inverse();
  for (uint32_t w = nwin - 1; w >= 1; w--) {
    F zi = inv * prefix[w];       // 1 / zzz_w
    inv = inv * pts[w].zzz;
    F zi2 = (zi * pts[w].zz).sqr();   // 1 / zz_w
    Affine<F> a{pts[w].x * zi2, pts[w].y * zi};
    a.store(out + uint64_t(w) * n + i);
  }
}

template <class F>
void msm_precompute_t(const void* d_in, void* d_out, uint64_t n, uint32_t nwin, uint32_t c, cudaStream_t s) {
  uint32_t blocks = uint32_t((n + 127) / 128);
  if (blocks) {
    msm_precompute_kernel<F><<<blocks, 128, 0, s>>>(reinterpret_cast<const Affine<F>*>(d_in),
                                                    reinterpret_cast<Affine<F>*>(d_out), n, nwin, c);
    count_launches(1);
  }
}

template <class F>
void msm_bases_to_mont_t(void* d_pts, uint64_t n, const uint32_t* b_mont, uint32_t* d_flag, cudaStream_t s) {
  uint32_t blocks = uint32_t((n + 127) / 128);
  if (blocks) {
    F b;
    static_assert(sizeof(F) % 4 == 0, "field element is a whole number of words");
    if (b_mont) memcpy(&b, b_mont, sizeof(F)); else memset(&b, 0, sizeof(F));
    msm_bases_to_mont_kernel<F><<<blocks, 128, 0, s>>>(reinterpret_cast<Affine<F>*>(d_pts), n, b,
                                                       b_mont ? d_flag : nullptr);
    count_launches(1);
  }
}
template <class F>
void msm_sum_affine_t(const void* d_pts, uint32_t n, void* d_out, cudaStream_t s) {
  msm_sum_affine_kernel<F><<<1, 32, 0, s>>>(reinterpret_cast<const Affine<F>*>(d_pts), n,
                                            reinterpret_cast<Affine<F>*>(d_out));
  count_launches(1);
}

}  // namespace zkb
