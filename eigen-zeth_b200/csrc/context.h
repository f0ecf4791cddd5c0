// context.h — the per-GPU context behind the C ABI (include/zkb.h): stream, scratch, cached
// NTT tables and MSM work buffers.  Internal to libzkb.so.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "../../include/zkb.h"
#include "comm.h"
#include "host_ff.hpp"
#include "msm.h"
#include "ntt.h"

namespace zkb {

// grow-only device allocation
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

struct NttDomain {          // everything that depends on (log_n, direction)
  uint32_t log_n = 0;
  std::vector<uint32_t> ks; // pass widths
  DevBuf tw_lo, tw_hi, tw_hi_ninv;
  uint32_t lo_bits = 0;
  std::vector<DevBuf> full; // full inter-pass twiddle tables, one per column pass (1/n folded into pass 0 of an inverse)
  bool full_ready = false;
};

struct CosetTables {        // per log_n
  DevBuf fwd;               // g^i
  DevBuf inv;               // g^{-i} / n
  DevBuf inv_z;             // g^{-i} / (n * (g^n - 1))     (Groth16 H pass)
  DevBuf ninv;              // constant 1/n (single-pass inverse)
};

struct MsmScratch {
  DevBuf buckets2, flag;
  // one proof = five MSMs whose latency-bound phases run side by side: partial lists per MSM and a
  // second set of tail scratch for the G2 reduction
  DevBuf multi_part, group_sums2, window_sums2;
  DevBuf keys_in, vals_in, keys_out, vals_out, sort_temp, m_valid, buckets, part1_keys,
      part1_pts, part2_keys, part2_pts, group_sums, window_sums, result_xyzz, result_affine;
};

}  // namespace zkb

struct zkb_bases {
  int device = 0;
  int is_g2 = 0;
  uint64_t n = 0;
  void* d_pts = nullptr;    // affine Montgomery, 64 / 128 B per point
  // precomputed window multiples: d_pts holds nwin blocks of n points, block w = 2^{c w} P_i
  int precomp_c = 0;
  uint32_t precomp_nwin = 0;
};

struct zkb_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  std::atomic_flag busy = ATOMIC_FLAG_INIT;
  int64_t msm_window = 0;
  int64_t ntt_full_twiddles = 1;           // n-entry inter-pass twiddle tables (1 modmul instead of 2)
  int64_t msm_profile = 0;
  int64_t msm_host_chunks = 0;             // 0 = auto
  int64_t pk_precompute = 1;               // proving-key queries carry precomputed window multiples
  int64_t pk_window = 0;                   // window width of proving-key queries (0 = by size, msm_auto_window)
  int64_t assemble_on_device = 0;          // 1: the one-CTA assembly kernel instead of the host ladders (comparison)
  cudaStream_t side[5] = {};               // latency-bound MSM phases of one proof, one stream per MSM
  cudaEvent_t side_ev[6] = {};
  cudaStream_t copy_stream = nullptr;      // H2D of scalar chunks, overlapped with compute
  cudaEvent_t copy_ev[16] = {};
  cudaEvent_t msm_ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  bool msm_ev_valid = false;

  zkb::DevBuf stage_fwd, stage_inv;        // 255-entry stage twiddle tables
  std::map<uint64_t, zkb::NttDomain*> domains;   // key = log_n * 2 + inverse
  std::map<uint32_t, zkb::CosetTables*> cosets;
  zkb::DevBuf ntt_tmp;                     // ping-pong buffer of the multi-pass NTT
  zkb::DevBuf io_a, io_b, io_c;            // staging for the host-buffer entry points
  zkb::MsmScratch msm;
  // multi-GPU (comm.cu / dist.cu)
  zkb::Comm* comm = nullptr;
  int64_t comm_force_staging = 0;          // test switch: NCCL send/recv staging even when peer mapping works
  std::map<std::string, zkb::DevBuf*> dist_tables;   // four-step twiddle / index-power tables
  zkb::DevBuf dist_small;                  // all-gather scratch of the distributed MSM / prove
};

namespace zkb {

// RAII guard implementing the "one call at a time per ctx" rule
struct BusyGuard {
  zkb_ctx* c;
  bool ok;
  explicit BusyGuard(zkb_ctx* ctx) : c(ctx), ok(!ctx->busy.test_and_set()) {}
  ~BusyGuard() { if (ok) c->busy.clear(); }
};

int set_error(zkb_ctx* ctx, int code, const std::string& msg);
int cuda_fail(zkb_ctx* ctx, cudaError_t e, const char* what);

#define ZKB_CUDA(ctx, expr)                                                   \
  do {                                                                        \
    cudaError_t _e = (expr);                                                  \
    if (_e != cudaSuccess) return ::zkb::cuda_fail((ctx), _e, #expr);         \
  } while (0)

// internal device-level entry points (used by the C ABI wrappers and by the Groth16 pipeline)
// `batch` contiguous transforms of 2^log_n elements each (tables, if given, cover the whole batch).
// skip_ninv: an inverse transform without the 1/n factor (the caller folds it elsewhere).
struct NttRemap {           // fused exchange of the four-step NTT, see NttPassArgs
  int in_ilv = 0;           // input batch is element-major (i * batch + b); batch must be a power of two
  int out_win = 0;          // output goes to per-rank windows instead of d_out
  uint32_t out_col_bits = 0;
  void* out_wins[16] = {};
};
int ntt_run(zkb_ctx* ctx, const void* d_in, void* d_out, uint32_t log_n, int inverse,
            const void* pre_scale, const void* post_scale, int skip_ninv, int mont_in,
            int mont_out, uint64_t batch = 1, const NttRemap* remap = nullptr);
int coset_tables(zkb_ctx* ctx, uint32_t log_n, CosetTables** out);
int msm_exec(zkb_ctx* ctx, const zkb_bases* b, uint64_t first, const void* d_scalars, uint64_t n,
             void* d_out_affine);
int msm_exec_multi(zkb_ctx* ctx, const zkb_bases* const* bs, int k, const void* d_scalars, uint64_t first,
                   uint64_t n, void* const* d_outs, cudaEvent_t* ev_opt);
int msm_exec_prove(zkb_ctx* ctx, const zkb_bases* a, const zkb_bases* b1, const zkb_bases* l, const zkb_bases* b2,
                   const zkb_bases* h, const void* d_z, uint64_t nz, const void* d_h, uint64_t nh, void* d_res,
                   cudaEvent_t* ev4_opt);
// host scalars: H2D in chunks on the copy stream, overlapped with the per-chunk MSM front half
int msm_exec_host(zkb_ctx* ctx, const zkb_bases* b, uint64_t first, const uint8_t* h_scalars,
                  uint64_t n, void* d_out_affine);
int bases_precompute_nolock(zkb_ctx* ctx, zkb_bases* b, int c);
int bases_to_mont_checked(zkb_ctx* ctx, zkb_bases* b);
int bases_from_scalars(zkb_ctx* ctx, const void* d_scalars, uint64_t n, int is_g2, void* d_out);
int bases_download_nolock(zkb_ctx* ctx, const zkb_bases* b, uint64_t first, uint64_t count,
                          uint8_t* out);

void host_fr_root(uint32_t log_n, int inverse, uint32_t out_mont[8]);
// proof assembly on the host (verify.cpp): fixed = alpha_g1 | beta_g1 | delta_g1 | beta_g2 | delta_g2 canonical
bool host_groth16_assemble(const uint8_t fixed[448], const uint8_t sums[ZKB_INTERMEDIATE_BYTES], const uint8_t r[32],
                           const uint8_t s[32], uint8_t proof_out[ZKB_PROOF_BYTES], const void* tables);
void* host_assemble_tables_create(const uint8_t fixed[448]);   // window tables of the key's delta_g1 / delta_g2
void host_assemble_tables_free(void* t);

// ---- distributed building blocks (dist.cu); collective over ctx->comm
// one four-step transform of this rank's layout-`d` block x (n / world elements, Montgomery) into
// layout log_n - d, in place; pre / post: optional element-wise tables in the input / output layout
int dist_ntt_transform(zkb_ctx* ctx, void* d_x, uint32_t log_n, uint32_t d, int inverse, const void* d_pre,
                       const void* d_post);
// table[pos] = scale * base^{global index of pos} for this rank's layout-d block (Montgomery words in)
int dist_index_power_table(zkb_ctx* ctx, uint32_t log_n, uint32_t d, const host::HFr& base,
                           const host::HFr& scale, const void** out);
// out = sum over the ranks of this rank's affine point (64 / 128 canonical bytes, device)
int dist_point_allreduce(zkb_ctx* ctx, const void* d_point, int is_g2, void* d_out);

}  // namespace zkb
