// msm_g2.cu — BN254 G2 instantiation of the MSM pipeline (coordinates in Fq2 on the twist).
// Three register-level forms of the same field, one memory layout (ff.cuh):
//   Fq2   every product inlined          -> the bucket walk (throughput-bound, best pipe use)
//   Fq2L  one element over a lane pair   -> the latency-bound kernels around the walk (partial merges,
//                                           bucket reduction, window sums, Horner): half the dependent
//                                           IMAD.WIDE chain per lane, twice the warps on idle SMs
//   Fq2C  out-of-line products           -> the same kernels without lane pairing (comparison), and
//                                           the setup kernels (small code)
// Process-wide comparison switch "msm_g2_pair": bit 0 = lane-pair WALK (measured 5 % slower than the
// one-thread walk), bit 1 = lane-pair latency-bound kernels (default).
#include "msm_impl.cuh"

namespace zkb {
template <> struct L1Cfg<Fq2C> { static constexpr int kMinBlocks = 2; };

void msm_accumulate_g2(const MsmPlan& p, const MsmBuffers& b, const void* d_bases, void* bucket_target,
                       cudaStream_t s) {
  const bool pair_walk = p.g2_pair & 1, pair_tail = p.g2_pair & 2;
  if (pair_walk && pair_tail) msm_accumulate_t<Fq2, Fq2L, Fq2L>(p, b, d_bases, bucket_target, s);
  else if (pair_walk) msm_accumulate_t<Fq2, Fq2L, Fq2C>(p, b, d_bases, bucket_target, s);
  else if (pair_tail) msm_accumulate_t<Fq2, Fq2, Fq2L>(p, b, d_bases, bucket_target, s);
  else msm_accumulate_t<Fq2, Fq2, Fq2C>(p, b, d_bases, bucket_target, s);
}
void msm_walk_g2(const MsmPlan& p, const MsmBuffers& b, const void* d_bases, void* bucket_target, cudaStream_t s) {
  if (p.g2_pair & 1) msm_walk_t<Fq2, Fq2L>(p, b, d_bases, bucket_target, s);
  else msm_walk_t<Fq2, Fq2>(p, b, d_bases, bucket_target, s);
}
void msm_post_g2(const MsmPlan& p, const MsmBuffers& b, void* bucket_target, cudaStream_t s) {
  if (p.g2_pair & 2) msm_post_t<Fq2, Fq2L>(p, b, bucket_target, s);
  else msm_post_t<Fq2, Fq2C>(p, b, bucket_target, s);
}
void msm_merge_g2(const MsmPlan& p, void* main_buckets, const void* part_buckets, cudaStream_t s) {
  msm_merge_t<Fq2C>(p, main_buckets, part_buckets, s);
}
void msm_tail_g2(const MsmPlan& p, const MsmBuffers& b, cudaStream_t s) {
  if (p.g2_pair & 2) msm_tail_t<Fq2L, Fq2C>(p, b, s);
  else msm_tail_t<Fq2C>(p, b, s);
}
void msm_bases_to_mont_g2(void* d_pts, uint64_t n, const uint32_t* b_mont, uint32_t* d_flag, cudaStream_t s) {
  msm_bases_to_mont_t<Fq2>(d_pts, n, b_mont, d_flag, s);
}
void msm_sum_affine_g2(const void* d_pts, uint32_t n, void* d_out, cudaStream_t s) {
  msm_sum_affine_t<Fq2C>(d_pts, n, d_out, s);
}
void msm_precompute_g2(const void* d_in, void* d_out, uint64_t n, uint32_t nwin, uint32_t c, cudaStream_t s) {
  msm_precompute_t<Fq2C>(d_in, d_out, n, nwin, c, s);
}
// chunks (not threads) walked concurrently per SM
uint32_t msm_l1_threads_per_sm_g2(int pair) {
  return (pair & 1) ? 128u * L1Cfg<Fq2L>::kMinBlocks / Fq2L::kLanes : 128u * L1Cfg<Fq2>::kMinBlocks;
}
}  // namespace zkb
