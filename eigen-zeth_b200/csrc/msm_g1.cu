// msm_g1.cu — BN254 G1 instantiation of the MSM pipeline (coordinates in Fq).
#include "msm_impl.cuh"

namespace zkb {
void msm_accumulate_g1(const MsmPlan& p, const MsmBuffers& b, const void* d_bases, void* bucket_target,
                       cudaStream_t s) {
  msm_accumulate_t<Fq>(p, b, d_bases, bucket_target, s);
}
void msm_walk_g1(const MsmPlan& p, const MsmBuffers& b, const void* d_bases, void* bucket_target, cudaStream_t s) {
  msm_walk_t<Fq>(p, b, d_bases, bucket_target, s);
}
void msm_post_g1(const MsmPlan& p, const MsmBuffers& b, void* bucket_target, cudaStream_t s) {
  msm_post_t<Fq>(p, b, bucket_target, s);
}
void msm_merge_g1(const MsmPlan& p, void* main_buckets, const void* part_buckets, cudaStream_t s) {
  msm_merge_t<Fq>(p, main_buckets, part_buckets, s);
}
void msm_tail_g1(const MsmPlan& p, const MsmBuffers& b, cudaStream_t s) { msm_tail_t<Fq>(p, b, s); }
void msm_tail_multi_g1(const MsmPlan& p, const MsmBuffers& b, void* const* d_outs, uint32_t n, cudaStream_t s) {
  MsmOutPtrs o{};
  o.n = n;
  for (uint32_t i = 0; i < n && i < 8; i++) o.p[i] = d_outs[i];
  msm_tail_t<Fq>(p, b, s, &o);
}
void msm_bases_to_mont_g1(void* d_pts, uint64_t n, const uint32_t* b_mont, uint32_t* d_flag, cudaStream_t s) {
  msm_bases_to_mont_t<Fq>(d_pts, n, b_mont, d_flag, s);
}
void msm_sum_affine_g1(const void* d_pts, uint32_t n, void* d_out, cudaStream_t s) {
  msm_sum_affine_t<Fq>(d_pts, n, d_out, s);
}
void msm_precompute_g1(const void* d_in, void* d_out, uint64_t n, uint32_t nwin, uint32_t c, cudaStream_t s) {
  msm_precompute_t<Fq>(d_in, d_out, n, nwin, c, s);
}
uint32_t msm_l1_threads_per_sm_g1() { return 128u * L1Cfg<Fq>::kMinBlocks; }
}  // namespace zkb
