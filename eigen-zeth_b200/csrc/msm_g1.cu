// msm_g1.cu — BN254 G1 instantiation of the MSM pipeline (coordinates in Fq).
#include "msm_impl.cuh"

namespace zkb {
void msm_run_g1(const MsmPlan& p, const MsmBuffers& b, const void* d_scalars, uint64_t n,
                const void* d_bases, cudaStream_t s) {
  msm_run_t<Fq>(p, b, d_scalars, n, d_bases, s);
}
void msm_bases_to_mont_g1(void* d_pts, uint64_t n, cudaStream_t s) { msm_bases_to_mont_t<Fq>(d_pts, n, s); }
void msm_sum_affine_g1(const void* d_pts, uint32_t n, void* d_out, cudaStream_t s) {
  msm_sum_affine_t<Fq>(d_pts, n, d_out, s);
}
}  // namespace zkb
