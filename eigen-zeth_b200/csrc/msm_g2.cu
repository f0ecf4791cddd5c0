// msm_g2.cu — BN254 G2 instantiation of the MSM pipeline (coordinates in Fq2 on the twist).
// The bucket walk runs on Fq2 (every product inlined); the latency-bound kernels around it run on
// Fq2C, the same field with out-of-line products (ff.cuh) — same memory layout.
#include "msm_impl.cuh"

namespace zkb {
template <> struct L1Cfg<Fq2C> { static constexpr int kMinBlocks = 2; };

void msm_accumulate_g2(const MsmPlan& p, const MsmBuffers& b, const void* d_bases, void* bucket_target,
                       cudaStream_t s) {
  // level-1 walk: one thread per point (8 warps / SM, 255 registers), or — comparison switch — one Fq2
  // point operation split over a lane pair (Fq2L, 16 warps / SM; 5 % slower on B200)
  if (p.g2_pair) msm_accumulate_t<Fq2C, Fq2L>(p, b, d_bases, bucket_target, s);
  else msm_accumulate_t<Fq2C, Fq2>(p, b, d_bases, bucket_target, s);
}
void msm_merge_g2(const MsmPlan& p, void* main_buckets, const void* part_buckets, cudaStream_t s) {
  msm_merge_t<Fq2C>(p, main_buckets, part_buckets, s);
}
void msm_tail_g2(const MsmPlan& p, const MsmBuffers& b, cudaStream_t s) { msm_tail_t<Fq2C>(p, b, s); }
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
  return pair ? 128u * L1Cfg<Fq2L>::kMinBlocks / Fq2L::kLanes : 128u * L1Cfg<Fq2>::kMinBlocks;
}
}  // namespace zkb
