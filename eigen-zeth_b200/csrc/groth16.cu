// groth16.cu — the Groth16 prover pipeline on one B200: R1CS matrix-vector products, the 7 NTTs
// of the QAP witness map, the five MSMs and proof assembly; plus synthetic proving-key generation.
//
// This is the device-side body of GenFinalProof (reference proto/prover/v1/prover.proto:130-148,
// requested from src/prover/provider.rs:472-483).  The algorithm is ark-groth16's
// create_proof_with_reduction_and_matrices with the LibsnarkReduction witness map (SURVEY.md
// Appendix B): domain = 2^log_n >= #constraints + #instance, coefficient-form H, coset shift 5.
#include <cuda_runtime.h>

#include <array>
#include <atomic>
#include <chrono>
#include <cstring>
#include <map>
#include <new>
#include <vector>

#include "context.h"
#include "ec.cuh"

using namespace zkb;
using zkb::host::HFr;

struct zkb_pk {
  int device = 0;
  uint32_t log_n = 0;
  uint64_t n = 0, m = 0, ncons = 0;
  zkb_bases *a_query = nullptr, *b_g1_query = nullptr, *b_g2_query = nullptr, *l_query = nullptr,
            *h_query = nullptr;
  // fixed points, affine Montgomery, device: alpha_g1 | beta_g1 | delta_g1 (3 x 64 B) and
  // beta_g2 | delta_g2 (2 x 128 B)
  void* d_fixed = nullptr;
  uint8_t h_fixed[448] = {};         // the same five points, canonical bytes, for the host-side assembly
  mutable void* h_tables = nullptr;  // fixed-base window tables of delta_g1 / delta_g2 (built at first use)
  // CSR matrices A, B, C (device)
  uint32_t* d_row_ptr[3] = {nullptr, nullptr, nullptr};
  uint32_t* d_col[3] = {nullptr, nullptr, nullptr};
  uint32_t* d_cidx[3] = {nullptr, nullptr, nullptr};
  void* d_coeff_table = nullptr;     // Fr Montgomery
  // work vectors (device, n x 32 B each) and witness staging
  void *d_a = nullptr, *d_b = nullptr, *d_c = nullptr, *d_z = nullptr, *d_zm = nullptr;
  void* d_results = nullptr;         // msm_a 64 | msm_b1 64 | msm_b2 128 | msm_l 64 | msm_h 64
  void* d_proof = nullptr;           // 256 B + r,s staging (64 B)
  // point-range shard held by this key: query q covers [qfirst[q], qfirst[q] + qcount[q]) of the
  // full query (q = 0 a, 1 b_g1, 2 b_g2, 3 l, 4 h); world == 1 means the whole key
  uint32_t n_inst = 2;       // instance variables incl. the constant one: z[0] = 1, z[1..n_inst)
  uint32_t rank = 0, world = 1;
  uint32_t h_low_bits = 0;   // != 0: h_query is held in "layout d" of the four-step NTT (csrc/dist.cu)
  // 1: snarkjs key (zkb_pk_desc.h_basis): h_query = L_{2i+1}(tau)/delta over the size-2n domain, n points;
  // A holds the input rows explicitly, there is no C matrix, d_shift = w_2n^i (Montgomery, n entries)
  uint32_t h_basis = 0;
  void* d_shift = nullptr;
  uint64_t qfirst[5] = {0, 0, 0, 0, 0}, qcount[5] = {0, 0, 0, 0, 0};
  // the prove entry points write the key's work vectors (d_a..d_zm, d_results, d_proof) behind a
  // const zkb_pk*: one prove at a time per key, whatever context it is called through
  mutable std::atomic_flag busy = ATOMIC_FLAG_INIT;
};

namespace zkb {
struct PkGuard {
  const zkb_pk* pk;
  bool ok;
  explicit PkGuard(const zkb_pk* k) : pk(k), ok(k && !k->busy.test_and_set()) {}
  ~PkGuard() { if (ok) pk->busy.clear(); }
};
}  // namespace zkb

namespace zkb {
int dist_sum_blocks(zkb_ctx* ctx, const void* d_block, uint32_t bytes, const uint32_t* offs, const uint32_t* g2s,
                    uint32_t n_slots, void* d_out);   // dist.cu

// ---------------------------------------------------------------- kernels
__global__ void fr_to_mont_copy_kernel(const Fr* in, Fr* out, uint64_t n) {
  uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  Fr::load(in + i).to_mont().store(out + i);
}

// out[j] = sum_k coeff[cidx[k]] * z[col[k]] over row j of a CSR matrix; rows >= nrows: see below
__global__ void r1cs_spmv_kernel(const uint32_t* __restrict__ row_ptr, const uint32_t* __restrict__ col,
                                 const uint32_t* __restrict__ cidx, const Fr* __restrict__ coeffs,
                                 const Fr* __restrict__ z, Fr* __restrict__ out, uint64_t nrows,
                                 uint64_t n, int copy_instance, uint32_t n_instance,
                                 uint32_t log_n, uint32_t low_bits, uint32_t lows, uint32_t low0) {
  uint64_t pos = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
  if (pos >= n) return;
  // low_bits == 0: out[j] for every row j < n.  Otherwise only this rank's block of "layout d"
  // (csrc/dist.cu): out[low_loc * H + high] = row (low0 + low_loc) + (high << low_bits), n = lows * H.
  uint64_t j = pos;
  if (low_bits) {
    const uint32_t hbits = log_n - low_bits;
    j = (low0 + (pos >> hbits)) + ((pos & ((1ull << hbits) - 1)) << low_bits);
  }
  Fr acc = Fr::zero();
  if (j < nrows) {
    uint32_t lo = row_ptr[j], hi = row_ptr[j + 1];
    for (uint32_t k = lo; k < hi; k++) {
      Fr c = Fr::load_nc(coeffs + cidx[k]);
      Fr v = Fr::load_nc(z + col[k]);
      acc = acc + c * v;
    }
  } else if (copy_instance && j - nrows < n_instance) {
    acc = Fr::load_nc(z + (j - nrows));   // ark-groth16 input-consistency rows of A
  }
  acc.store(out + pos);
}

// a[i] = a[i] * b[i] - c[i]    (the 1/Z(g) factor is folded into the last coset-iNTT table)
__global__ void qap_pointwise_kernel(Fr* a, const Fr* b, const Fr* c, uint64_t n) {
  uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  (Fr::load(a + i) * Fr::load(b + i) - Fr::load(c + i)).store(a + i);
}

// snarkjs keys: c[i] = a[i] * b[i] on the domain (no C matrix), and the H scalars
// a[i] = a[i] * b[i] - c[i] on the odd coset, stored as canonical integers for the MSM
__global__ void qap_product_kernel(const Fr* a, const Fr* b, Fr* c, uint64_t n) {
  uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  (Fr::load(a + i) * Fr::load(b + i)).store(c + i);
}
__global__ void qap_pointwise_canonical_kernel(Fr* a, const Fr* b, const Fr* c, uint64_t n) {
  uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  (Fr::load(a + i) * Fr::load(b + i) - Fr::load(c + i)).from_mont().store(a + i);
}

template <class F>
__device__ Affine<F> load_canonical_affine(const void* p) {
  Affine<F> a = Affine<F>::load(p);
  a.x = a.x.to_mont();
  a.y = a.y.to_mont();
  return a;
}
template <class F>
__device__ void store_canonical_affine(const XYZZ<F>& p, void* out) {
  Affine<F> a = p.to_affine();
  a.x = a.x.from_mont();
  a.y = a.y.from_mont();
  a.store(out);
}

// Proof assembly (K15).  One CTA, four warps; lane 0 of each warp works:
//  phase 1: w0: r*delta1   w1: s*delta1   w2: s*delta2   w3: (r*s)*delta1
//  phase 2: w0: A = alpha + msm_a + r*delta1 ; s*A      w1: B1 = beta1 + msm_b1 + s*delta1 ; r*B1
//           w2: B2 = beta2 + msm_b2 + s*delta2
//  phase 3: w0: C = msm_l + msm_h + s*A + r*B1 - rs*delta1
// results: canonical affine; fixed: Montgomery affine; rs: r | s canonical; proof: A | B2 | C.
__global__ void groth16_assemble_kernel(const uint8_t* results, const uint8_t* fixed,
                                        const uint32_t* rs, uint8_t* proof) {
  __shared__ uint4 sm_raw[(4 * sizeof(G1XYZZ) + sizeof(G2XYZZ)) / 16];
  G1XYZZ* s1 = reinterpret_cast<G1XYZZ*>(sm_raw);            // [0] r*d1 -> s*A, [1] s*d1 -> r*B1, [2] rs*d1
  G2XYZZ* s2 = reinterpret_cast<G2XYZZ*>(s1 + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t r[8], s[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { r[i] = rs[i]; s[i] = rs[8 + i]; }
  const G1Affine alpha = G1Affine::load(fixed), beta1 = G1Affine::load(fixed + 64),
                 delta1 = G1Affine::load(fixed + 128);
  if (lane == 0) {
    if (warp == 0) mul_scalar(G1XYZZ::from_affine(delta1), r).store(s1 + 0);
    if (warp == 1) mul_scalar(G1XYZZ::from_affine(delta1), s).store(s1 + 1);
    if (warp == 2) {
      G2Affine delta2 = G2Affine::load(fixed + 192 + 128);
      mul_scalar(G2XYZZ::from_affine(delta2), s).store(s2);
    }
    if (warp == 3) {
      Fr rm, sm;
#pragma unroll
      for (int i = 0; i < 8; i++) { rm.v[i] = r[i]; sm.v[i] = s[i]; }
      Fr prod = (rm.to_mont() * sm.to_mont()).from_mont();
      mul_scalar(G1XYZZ::from_affine(delta1), prod.v).store(s1 + 2);
    }
  }
  __syncthreads();
  if (lane == 0) {
    if (warp == 0) {
      G1XYZZ A = G1XYZZ::load(s1 + 0);
      A.madd(alpha);
      A.madd(load_canonical_affine<Fq>(results));
      G1Affine Aa = A.to_affine();
      G1Affine out = Aa;
      out.x = out.x.from_mont();
      out.y = out.y.from_mont();
      out.store(proof);
      mul_scalar(G1XYZZ::from_affine(Aa), s).store(s1 + 0);
    }
    if (warp == 1) {
      G1XYZZ B1 = G1XYZZ::load(s1 + 1);
      B1.madd(beta1);
      B1.madd(load_canonical_affine<Fq>(results + 64));
      G1Affine Ba = B1.to_affine();
      mul_scalar(G1XYZZ::from_affine(Ba), r).store(s1 + 1);
    }
    if (warp == 2) {
      G2XYZZ B2 = G2XYZZ::load(s2);
      B2.madd(G2Affine::load(fixed + 192));
      B2.madd(load_canonical_affine<Fq2>(results + 128));
      store_canonical_affine<Fq2>(B2, proof + 64);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    G1XYZZ Cc = G1XYZZ::load(s1 + 0);
    Cc.add(G1XYZZ::load(s1 + 1));
    Cc.add(G1XYZZ::load(s1 + 2).neg());
    Cc.madd(load_canonical_affine<Fq>(results + 256));
    Cc.madd(load_canonical_affine<Fq>(results + 320));
    store_canonical_affine<Fq>(Cc, proof + 192);
  }
}

// ---------------------------------------------------------------- host helpers
static void hfr_words(const HFr& a, uint32_t out[8]) { std::memcpy(out, a.l, 32); }

struct HostR1CS {
  uint64_t m, ncons;
  std::vector<uint32_t> row_ptr[3], col[3], cidx[3];
  std::vector<uint32_t> coeff_values;   // small integer coefficients, index = cidx
};

// The synthetic chain circuit: constraint j <= m-5: (z[j+2] + (j%7+1)) * z[j+3] = z[j+4];
// then z[m-1] * z[m-2] = z[1]; then (z[1] + z[2]) * 1 = z[1] + z[2].
static HostR1CS synthetic_r1cs(uint32_t log_n) {
  HostR1CS cs;
  cs.m = 1ull << log_n;
  const uint64_t m = cs.m;
  for (uint32_t v = 0; v <= 7; v++) cs.coeff_values.push_back(v);   // cidx == coefficient value
  auto push = [&](int mat, std::initializer_list<std::pair<uint32_t, uint32_t>> row) {
    for (auto& e : row) {
      cs.col[mat].push_back(e.first);
      cs.cidx[mat].push_back(e.second);
    }
    cs.row_ptr[mat].push_back(uint32_t(cs.col[mat].size()));
  };
  for (int k = 0; k < 3; k++) cs.row_ptr[k].push_back(0);
  for (uint64_t j = 0; j + 4 < m; j++) {
    push(0, {{0u, uint32_t(j % 7 + 1)}, {uint32_t(j + 2), 1u}});
    push(1, {{uint32_t(j + 3), 1u}});
    push(2, {{uint32_t(j + 4), 1u}});
  }
  push(0, {{uint32_t(m - 1), 1u}});
  push(1, {{uint32_t(m - 2), 1u}});
  push(2, {{1u, 1u}});
  push(0, {{1u, 1u}, {2u, 1u}});
  push(1, {{0u, 1u}});
  push(2, {{1u, 1u}, {2u, 1u}});
  cs.ncons = cs.row_ptr[0].size() - 1;
  return cs;
}

static HFr hpow_u64(HFr b, uint64_t e) {
  HFr r = HFr::one();
  while (e) {
    if (e & 1) r = r * b;
    b = b.sqr();
    e >>= 1;
  }
  return r;
}

template <class T>
static cudaError_t upload_vec(const std::vector<T>& v, T** d, cudaStream_t s) {
  cudaError_t e = cudaMalloc(d, (v.size() ? v.size() : 1) * sizeof(T));
  if (e != cudaSuccess) return e;
  return cudaMemcpyAsync(*d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s);
}

static int make_query(zkb_ctx* ctx, const std::vector<uint64_t>& scalars_canon, uint64_t first,
                      uint64_t count, int is_g2, zkb_bases** out) {
  zkb_bases* b = new (std::nothrow) zkb_bases();
  if (!b) return set_error(ctx, ZKB_ENOMEM, "out of host memory");
  b->device = ctx->device;
  b->is_g2 = is_g2;
  b->n = count;
  cudaError_t e = cudaMalloc(&b->d_pts, (count ? count : 1) * (is_g2 ? 128 : 64));
  if (e != cudaSuccess) {
    delete b;
    return cuda_fail(ctx, e, "query allocation");
  }
  *out = b;
  ZKB_CUDA(ctx, ctx->io_a.ensure(count * 32 + 32));
  ZKB_CUDA(ctx, cudaMemcpyAsync(ctx->io_a.p, scalars_canon.data() + 4 * first, count * 32,
                                cudaMemcpyHostToDevice, ctx->stream));
  int rc = bases_from_scalars(ctx, ctx->io_a.p, count, is_g2, b->d_pts);
  if (rc) return rc;
  ZKB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // scalars_canon may die after return
  return ZKB_OK;
}

// CSR matrices + coefficient table -> device, and the per-key work buffers
static int pk_upload_circuit(zkb_ctx* ctx, zkb_pk* pk, const HostR1CS& cs, const std::vector<HFr>& coeff_m) {
  cudaError_t e;
  for (int k = 0; k < 3; k++) {
    if ((e = upload_vec(cs.row_ptr[k], &pk->d_row_ptr[k], ctx->stream)) != cudaSuccess ||
        (e = upload_vec(cs.col[k], &pk->d_col[k], ctx->stream)) != cudaSuccess ||
        (e = upload_vec(cs.cidx[k], &pk->d_cidx[k], ctx->stream)) != cudaSuccess)
      return cuda_fail(ctx, e, "matrix upload");
  }
  std::vector<uint32_t> ctab(8 * (coeff_m.size() ? coeff_m.size() : 1));
  for (size_t i = 0; i < coeff_m.size(); i++) hfr_words(coeff_m[i], &ctab[8 * i]);
  uint32_t* d_ct = nullptr;
  if ((e = upload_vec(ctab, &d_ct, ctx->stream)) != cudaSuccess) return cuda_fail(ctx, e, "coeff upload");
  pk->d_coeff_table = d_ct;
  const uint64_t len = pk->n > pk->m ? pk->n : pk->m;
  void** work[] = {&pk->d_a, &pk->d_b, &pk->d_c, &pk->d_z, &pk->d_zm};
  for (void** wp : work)
    if ((e = cudaMalloc(wp, len * 32)) != cudaSuccess) return cuda_fail(ctx, e, "work vectors");
  if ((e = cudaMalloc(&pk->d_results, ZKB_INTERMEDIATE_BYTES)) != cudaSuccess ||
      (e = cudaMalloc(&pk->d_proof, ZKB_PROOF_BYTES + 64)) != cudaSuccess ||
      (e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess)
    return cuda_fail(ctx, e, "result buffers");
  return ZKB_OK;
}

static void canon_into(const HFr& a, std::vector<uint64_t>& v, uint64_t i) {
  HFr c = a.from_mont();
  std::memcpy(&v[4 * i], c.l, 32);
}

}  // namespace zkb

extern "C" {

void zkb_pk_free(zkb_pk* pk) {
  if (!pk) return;
  int cur = 0;
  cudaGetDevice(&cur);
  cudaSetDevice(pk->device);
  zkb_bases_free(pk->a_query);
  zkb_bases_free(pk->b_g1_query);
  zkb_bases_free(pk->b_g2_query);
  zkb_bases_free(pk->l_query);
  zkb_bases_free(pk->h_query);
  void* ptrs[] = {pk->d_fixed, pk->d_coeff_table, pk->d_a, pk->d_b, pk->d_c, pk->d_z, pk->d_zm,
                  pk->d_results, pk->d_proof, pk->d_shift};
  for (void* p : ptrs) if (p) cudaFree(p);
  for (int k = 0; k < 3; k++) {
    if (pk->d_row_ptr[k]) cudaFree(pk->d_row_ptr[k]);
    if (pk->d_col[k]) cudaFree(pk->d_col[k]);
    if (pk->d_cidx[k]) cudaFree(pk->d_cidx[k]);
  }
  cudaSetDevice(cur);
  host_assemble_tables_free(pk->h_tables);
  delete pk;
}

uint64_t zkb_pk_n_vars(const zkb_pk* pk) { return pk ? pk->m : 0; }
uint32_t zkb_pk_log_n(const zkb_pk* pk) { return pk ? pk->log_n : 0; }

const zkb_bases* zkb_pk_query(const zkb_pk* pk, int which) {
  if (!pk) return nullptr;
  switch (which) {
    case 0: return pk->a_query;
    case 1: return pk->b_g1_query;
    case 2: return pk->b_g2_query;
    case 3: return pk->l_query;
    case 4: return pk->h_query;
    default: return nullptr;
  }
}

int zkb_synthetic_witness(uint32_t log_n, const uint8_t w0[32], const uint8_t w1[32],
                          uint8_t* z_out) {
  if (log_n < 3 || log_n > 28 || !w0 || !w1 || !z_out) return ZKB_EINVAL;
  const uint64_t m = 1ull << log_n;
  HFr a, b;
  if (!HFr::from_bytes(w0, &a) || !HFr::from_bytes(w1, &b)) return ZKB_EINVAL;
  HFr small[8];
  for (int i = 0; i < 8; i++) small[i] = HFr::from_u64(i);
  // z[j+4] = (z[j+2] + (j%7+1)) * z[j+3]: a two-term recurrence, streamed straight into the output
  HFr::one().to_bytes(z_out);
  a.to_bytes(z_out + 64);
  b.to_bytes(z_out + 96);
  HFr p2 = a, p1 = b;                    // z[j+2], z[j+3]
  for (uint64_t j = 0; j + 4 < m; j++) {
    HFr nx = (p2 + small[j % 7 + 1]) * p1;
    nx.to_bytes(z_out + 32 * (j + 4));
    p2 = p1;
    p1 = nx;
  }
  (p1 * p2).to_bytes(z_out + 32);        // z[1] = z[m-1] * z[m-2]
  return ZKB_OK;
}

int zkb_pk_synthetic(zkb_ctx* ctx, uint32_t log_n, uint64_t seed, zkb_pk** out,
                     uint8_t vk_out[ZKB_VK_BYTES]) {
  return zkb_pk_synthetic_shard(ctx, log_n, seed, 0, 1, out, vk_out);
}

int zkb_pk_synthetic_shard(zkb_ctx* ctx, uint32_t log_n, uint64_t seed, uint32_t rank, uint32_t world,
                           zkb_pk** out, uint8_t vk_out[ZKB_VK_BYTES]) {
  return zkb_pk_synthetic_shard_ex(ctx, log_n, seed, rank, world, 0, out, vk_out);
}

int zkb_pk_synthetic_shard_ex(zkb_ctx* ctx, uint32_t log_n, uint64_t seed, uint32_t rank, uint32_t world,
                              uint32_t h_low_bits, zkb_pk** out, uint8_t vk_out[ZKB_VK_BYTES]) {
  if (!ctx) return ZKB_EINVAL;
  BusyGuard guard(ctx);
  if (!guard.ok) return ZKB_EBUSY;
  if (!out || !vk_out) return set_error(ctx, ZKB_EINVAL, "null argument");
  if (log_n < 3 || log_n > 26) return set_error(ctx, ZKB_EINVAL, "log_n must be in [3, 26]");
  if (world == 0 || rank >= world || world > 64) return set_error(ctx, ZKB_EINVAL, "bad rank / world");
  if (h_low_bits && (h_low_bits >= log_n || ((1ull << h_low_bits) % world) || (world & (world - 1))))
    return set_error(ctx, ZKB_EINVAL, "bad h layout");
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  zkb_pk* pk = new (std::nothrow) zkb_pk();
  if (!pk) return set_error(ctx, ZKB_ENOMEM, "out of host memory");
  int rc = ZKB_OK;
  try {
    HostR1CS cs = synthetic_r1cs(log_n);
    const uint64_t n = 1ull << log_n, m = cs.m;
    pk->device = ctx->device;
    pk->log_n = log_n;
    pk->n = n;
    pk->m = m;
    pk->ncons = cs.ncons;
    pk->rank = rank;
    pk->world = world;
    pk->h_low_bits = h_low_bits;
    const uint64_t qlen[5] = {m, m, m, m, n - 1};   // l_query is stored padded to m (see below)
    for (int q = 0; q < 5; q++) {
      uint64_t base = qlen[q] / world, rem = qlen[q] % world;
      pk->qfirst[q] = rank * base + (rank < rem ? rank : rem);
      pk->qcount[q] = base + (rank < rem ? 1 : 0);
    }
    host::Xoshiro256ss rng(seed);
    HFr tau = rng.fr(), alpha = rng.fr(), beta = rng.fr(), gamma = rng.fr(), delta = rng.fr();
    // Lagrange basis at tau
    uint32_t wroot[8];
    host_fr_root(log_n, 0, wroot);
    HFr w;
    std::memcpy(w.l, wroot, 32);
    HFr zt = hpow_u64(tau, n) - HFr::one();
    HFr ninv = hpow_u64(HFr::from_u64(2), log_n).inverse();
    // u_i(tau), v_i(tau), w_i(tau) for the variables this key needs: the rank's range [lo, lo+cnt)
    // and the instance variables 0, 1 (kept in their own slots).  The Lagrange values
    // L_j = (tau^n - 1)/n * w^j / (tau - w^j) are streamed in blocks (one inversion per block), so
    // host memory stays O(cnt) whatever the circuit size.
    const uint64_t lo = pk->qfirst[0], cnt = pk->qcount[0];
    std::vector<HFr> uvw[3];
    for (auto& x : uvw) x.assign(cnt, HFr::zero());
    HFr inst[3][2];
    for (auto& x : inst) x[0] = x[1] = HFr::zero();
    std::vector<HFr> coeff_m;
    for (uint32_t c : cs.coeff_values) coeff_m.push_back(HFr::from_u64(c));
    {
      const uint64_t kBlock = 1ull << 16;
      std::vector<HFr> L(kBlock), den(kBlock), pre(kBlock);
      const HFr scale = zt * ninv;
      HFr wj = HFr::one();
      for (uint64_t j0 = 0; j0 < n; j0 += kBlock) {
        const uint64_t len = n - j0 < kBlock ? n - j0 : kBlock;
        HFr acc = HFr::one();
        for (uint64_t t = 0; t < len; t++) {
          den[t] = tau - wj;
          L[t] = scale * wj;
          wj = wj * w;
          pre[t] = acc;
          acc = acc * den[t];
        }
        HFr ia = acc.inverse();
        for (uint64_t t = len; t-- > 0;) {
          L[t] = L[t] * (ia * pre[t]);
          ia = ia * den[t];
        }
        for (uint64_t t = 0; t < len; t++) {
          const uint64_t j = j0 + t;
          if (j < cs.ncons) {
            for (int k = 0; k < 3; k++)
              for (uint32_t e2 = cs.row_ptr[k][j]; e2 < cs.row_ptr[k][j + 1]; e2++) {
                const uint64_t col = cs.col[k][e2];
                HFr term = coeff_m[cs.cidx[k][e2]] * L[t];
                if (col < 2) inst[k][col] = inst[k][col] + term;
                if (col >= lo && col < lo + cnt) uvw[k][col - lo] = uvw[k][col - lo] + term;
              }
          } else if (j - cs.ncons < 2) {      // ark-groth16 input-consistency rows of A
            const uint64_t col = j - cs.ncons;
            inst[0][col] = inst[0][col] + L[t];
            if (col >= lo && col < lo + cnt) uvw[0][col - lo] = uvw[0][col - lo] + L[t];
          }
        }
      }
    }
    HFr gi = gamma.inverse(), di = delta.inverse();
    // scalars (canonical) for every query; one reusable staging vector of the shard size
    uint64_t stage_len = cnt;
    if (pk->qcount[4] > stage_len) stage_len = pk->qcount[4];
    if (h_low_bits && (n / world) > stage_len) stage_len = n / world;
    std::vector<uint64_t> sc(4 * stage_len);
    for (uint64_t i = 0; i < cnt; i++) canon_into(uvw[0][i], sc, i);
    if ((rc = make_query(ctx, sc, 0, cnt, 0, &pk->a_query))) throw rc;
    for (uint64_t i = 0; i < cnt; i++) canon_into(uvw[1][i], sc, i);
    if ((rc = make_query(ctx, sc, 0, cnt, 0, &pk->b_g1_query))) throw rc;
    if ((rc = make_query(ctx, sc, 0, cnt, 1, &pk->b_g2_query))) throw rc;
    HFr abc[2];
    for (int i = 0; i < 2; i++) abc[i] = (beta * inst[0][i] + alpha * inst[1][i] + inst[2][i]) * gi;
    for (uint64_t i = 0; i < cnt; i++) {
      // l_query is kept at length m with the instance entries set to the point at infinity
      // (scalar 0), so that the A / B1 / B2 / L MSMs share one digit extraction and one sort
      if (lo + i < 2) canon_into(HFr::zero(), sc, i);
      else canon_into((beta * uvw[0][i] + alpha * uvw[1][i] + uvw[2][i]) * di, sc, i);
    }
    if ((rc = make_query(ctx, sc, 0, cnt, 0, &pk->l_query))) throw rc;
    for (auto& x : uvw) std::vector<HFr>().swap(x);
    // h_query[k] = tau^k (tau^n - 1) / delta
    const HFr h0 = zt * di;
    if (h_low_bits) {
      // in the output layout of the distributed NTT: local [low_loc][high] holds the base of
      // coefficient (rank*lows + low_loc) + (high << h_low_bits); index n-1 (h has n-1
      // coefficients) is the point at infinity
      const uint64_t lows = (1ull << h_low_bits) / world, highs = n >> h_low_bits;
      const HFr step = hpow_u64(tau, 1ull << h_low_bits);
      for (uint64_t l2 = 0; l2 < lows; l2++) {
        HFr tp = h0 * hpow_u64(tau, rank * lows + l2);
        for (uint64_t hi = 0; hi < highs; hi++) {
          const uint64_t k = (rank * lows + l2) + (hi << h_low_bits);
          canon_into(k == n - 1 ? HFr::zero() : tp, sc, l2 * highs + hi);
          tp = tp * step;
        }
      }
      pk->qfirst[4] = 0;
      pk->qcount[4] = lows * highs;
    } else {
      HFr tp = h0 * hpow_u64(tau, pk->qfirst[4]);
      for (uint64_t i = 0; i < pk->qcount[4]; i++) {
        canon_into(tp, sc, i);
        tp = tp * tau;
      }
    }
    if ((rc = make_query(ctx, sc, 0, pk->qcount[4], 0, &pk->h_query))) throw rc;
    std::vector<uint64_t>().swap(sc);
    if (ctx->pk_precompute) {
      zkb_bases* qs[5] = {pk->a_query, pk->b_g1_query, pk->b_g2_query, pk->l_query, pk->h_query};
      for (zkb_bases* q : qs)
        if (q->n && (rc = bases_precompute_nolock(ctx, q, int(ctx->pk_window)))) throw rc;
    }
    // fixed points: [alpha, beta, delta, ic0, ic1] * G1 and [beta, delta, gamma] * G2
    std::vector<uint64_t> f1(4 * 5), f2(4 * 3);
    canon_into(alpha, f1, 0); canon_into(beta, f1, 1); canon_into(delta, f1, 2);
    canon_into(abc[0], f1, 3); canon_into(abc[1], f1, 4);
    canon_into(beta, f2, 0); canon_into(delta, f2, 1); canon_into(gamma, f2, 2);
    zkb_bases *b1 = nullptr, *b2 = nullptr;
    if ((rc = make_query(ctx, f1, 0, 5, 0, &b1))) throw rc;
    if ((rc = make_query(ctx, f2, 0, 3, 1, &b2))) { zkb_bases_free(b1); throw rc; }
    cudaError_t e = cudaMalloc(&pk->d_fixed, 3 * 64 + 2 * 128);
    if (e == cudaSuccess) e = cudaMemcpyAsync(pk->d_fixed, b1->d_pts, 3 * 64, cudaMemcpyDeviceToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync((char*)pk->d_fixed + 192, b2->d_pts, 2 * 128, cudaMemcpyDeviceToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    // vk (canonical bytes): alpha_g1 | beta_g2 | gamma_g2 | delta_g2 | ic0 | ic1
    uint8_t g1c[5 * 64], g2c[3 * 128];
    int r1 = bases_download_nolock(ctx, b1, 0, 5, g1c);
    int r2 = bases_download_nolock(ctx, b2, 0, 3, g2c);
    zkb_bases_free(b1);
    zkb_bases_free(b2);
    if (e != cudaSuccess) throw cuda_fail(ctx, e, "fixed points");
    if (r1 || r2) throw (r1 ? r1 : r2);
    std::memcpy(pk->h_fixed, g1c, 192);
    std::memcpy(pk->h_fixed + 192, g2c, 256);
    std::memcpy(vk_out, g1c, 64);
    std::memcpy(vk_out + 64, g2c, 128);
    std::memcpy(vk_out + 192, g2c + 256, 128);
    std::memcpy(vk_out + 320, g2c + 128, 128);
    std::memcpy(vk_out + 448, g1c + 192, 128);
    if ((rc = pk_upload_circuit(ctx, pk, cs, coeff_m))) throw rc;
  } catch (int code) {
    zkb_pk_free(pk);
    return code;
  } catch (const std::bad_alloc&) {
    zkb_pk_free(pk);
    return set_error(ctx, ZKB_ENOMEM, "out of host memory");
  }
  *out = pk;
  return ZKB_OK;
}

}  // extern "C" (reopened below)

// external bases (canonical affine bytes) -> device-resident zkb_bases, optionally zero-padded in front
static int bases_from_bytes(zkb_ctx* ctx, const uint8_t* pts, uint64_t count, uint64_t pad_front, int is_g2,
                            zkb_bases** out) {
  const size_t sz = is_g2 ? 128 : 64;
  for (uint64_t i = 0; i < count * (sz / 32); i++) {
    uint64_t w[4];
    std::memcpy(w, pts + 32 * i, 32);
    if (!host::HFq::lt_mod(w)) return set_error(ctx, ZKB_EINVAL, "non-canonical Fq coordinate in a query");
  }
  zkb_bases* b = new (std::nothrow) zkb_bases();
  if (!b) return set_error(ctx, ZKB_ENOMEM, "out of host memory");
  b->device = ctx->device;
  b->is_g2 = is_g2;
  b->n = count + pad_front;
  cudaError_t e = cudaMalloc(&b->d_pts, (b->n ? b->n : 1) * sz);
  if (e != cudaSuccess) {
    delete b;
    return cuda_fail(ctx, e, "query allocation");
  }
  *out = b;
  ZKB_CUDA(ctx, cudaMemsetAsync(b->d_pts, 0, pad_front * sz, ctx->stream));
  ZKB_CUDA(ctx, cudaMemcpyAsync((char*)b->d_pts + pad_front * sz, pts, count * sz, cudaMemcpyHostToDevice, ctx->stream));
  int rc = bases_to_mont_checked(ctx, b);
  if (rc) {
    zkb_bases_free(b);
    *out = nullptr;
  }
  return rc;
}

extern "C" {

int zkb_pk_load(zkb_ctx* ctx, const zkb_pk_desc* d, zkb_pk** out) {
  if (!ctx) return ZKB_EINVAL;
  BusyGuard guard(ctx);
  if (!guard.ok) return ZKB_EBUSY;
  if (!d || !out) return set_error(ctx, ZKB_EINVAL, "null argument");
  const uint64_t n = 1ull << d->log_n, m = d->n_vars, n_inst = uint64_t(d->n_public) + 1;
  if (d->h_basis > 1) return set_error(ctx, ZKB_EINVAL, "unknown h_basis");
  const bool snark = d->h_basis == 1;       // A already carries the input rows, C may be absent
  if (d->log_n < 1 || d->log_n > 26 || m < n_inst || m >= (1ull << 31) ||
      d->n_constraints + (snark ? 0 : n_inst) > n || n_inst > 64)
    return set_error(ctx, ZKB_EINVAL, "inconsistent circuit dimensions (domain must hold constraints + instance)");
  const int n_mat = (snark && !d->row_ptr[2]) ? 2 : 3;
  for (int k = 0; k < n_mat; k++)
    if (!d->row_ptr[k] || (d->row_ptr[k][d->n_constraints] && (!d->col[k] || !d->coeff[k])))
      return set_error(ctx, ZKB_EINVAL, "null matrix");
  if (!d->a_query || !d->b_g1_query || !d->b_g2_query || !d->h_query || (m > n_inst && !d->l_query) ||
      !d->alpha_g1 || !d->beta_g1 || !d->delta_g1 || !d->beta_g2 || !d->delta_g2)
    return set_error(ctx, ZKB_EINVAL, "null query");
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  zkb_pk* pk = new (std::nothrow) zkb_pk();
  if (!pk) return set_error(ctx, ZKB_ENOMEM, "out of host memory");
  int rc = ZKB_OK;
  try {
    pk->device = ctx->device;
    pk->log_n = d->log_n;
    pk->n = n;
    pk->m = m;
    pk->ncons = d->n_constraints;
    pk->n_inst = uint32_t(n_inst);
    pk->h_basis = d->h_basis;
    const uint64_t qlen[5] = {m, m, m, m, snark ? n : n - 1};
    for (int q = 0; q < 5; q++) pk->qcount[q] = qlen[q];
    // matrices: validate, de-duplicate coefficients into a table
    HostR1CS cs;
    cs.m = m;
    cs.ncons = d->n_constraints;
    std::vector<HFr> coeff_m;
    std::map<std::array<uint64_t, 4>, uint32_t> seen;
    for (int k = 0; k < 3; k++) {
      if (k >= n_mat) {                      // no C matrix: an empty one keeps the uploads uniform
        cs.row_ptr[k].assign(cs.ncons + 1, 0);
        continue;
      }
      const uint32_t nnz = d->row_ptr[k][cs.ncons];
      cs.row_ptr[k].assign(d->row_ptr[k], d->row_ptr[k] + cs.ncons + 1);
      cs.col[k].assign(d->col[k], d->col[k] + nnz);
      cs.cidx[k].resize(nnz);
      if (cs.row_ptr[k][0] != 0) throw set_error(ctx, ZKB_EINVAL, "row_ptr must start at 0");
      for (uint64_t j = 0; j < cs.ncons; j++)
        if (cs.row_ptr[k][j] > cs.row_ptr[k][j + 1]) throw set_error(ctx, ZKB_EINVAL, "row_ptr not monotone");
      for (uint32_t e2 = 0; e2 < nnz; e2++) {
        if (cs.col[k][e2] >= m) throw set_error(ctx, ZKB_EINVAL, "column index out of range");
        std::array<uint64_t, 4> key;
        std::memcpy(key.data(), d->coeff[k] + 32 * uint64_t(e2), 32);
        auto it = seen.find(key);
        if (it == seen.end()) {
          HFr c;
          if (!HFr::from_bytes(d->coeff[k] + 32 * uint64_t(e2), &c))
            throw set_error(ctx, ZKB_EINVAL, "non-canonical matrix coefficient");
          it = seen.emplace(key, uint32_t(coeff_m.size())).first;
          coeff_m.push_back(c);
        }
        cs.cidx[k][e2] = it->second;
      }
    }
    if ((rc = bases_from_bytes(ctx, d->a_query, m, 0, 0, &pk->a_query))) throw rc;
    if ((rc = bases_from_bytes(ctx, d->b_g1_query, m, 0, 0, &pk->b_g1_query))) throw rc;
    if ((rc = bases_from_bytes(ctx, d->b_g2_query, m, 0, 1, &pk->b_g2_query))) throw rc;
    // l_query arrives without the instance entries (ark-groth16); stored padded, see zkb_pk_query
    if ((rc = bases_from_bytes(ctx, d->l_query, m - n_inst, n_inst, 0, &pk->l_query))) throw rc;
    if ((rc = bases_from_bytes(ctx, d->h_query, snark ? n : n - 1, 0, 0, &pk->h_query))) throw rc;
    if (ctx->pk_precompute) {
      zkb_bases* qs[5] = {pk->a_query, pk->b_g1_query, pk->b_g2_query, pk->l_query, pk->h_query};
      for (zkb_bases* q : qs)
        if (q->n && (rc = bases_precompute_nolock(ctx, q, int(ctx->pk_window)))) throw rc;
    }
    // fixed points -> Montgomery on device: alpha_g1 | beta_g1 | delta_g1 | beta_g2 | delta_g2
    uint8_t f1[192], f2[256];
    std::memcpy(f1, d->alpha_g1, 64); std::memcpy(f1 + 64, d->beta_g1, 64); std::memcpy(f1 + 128, d->delta_g1, 64);
    std::memcpy(f2, d->beta_g2, 128); std::memcpy(f2 + 128, d->delta_g2, 128);
    std::memcpy(pk->h_fixed, f1, 192);
    std::memcpy(pk->h_fixed + 192, f2, 256);
    zkb_bases *b1 = nullptr, *b2 = nullptr;
    if ((rc = bases_from_bytes(ctx, f1, 3, 0, 0, &b1))) throw rc;
    if ((rc = bases_from_bytes(ctx, f2, 2, 0, 1, &b2))) { zkb_bases_free(b1); throw rc; }
    cudaError_t e = cudaMalloc(&pk->d_fixed, 3 * 64 + 2 * 128);
    if (e == cudaSuccess) e = cudaMemcpyAsync(pk->d_fixed, b1->d_pts, 3 * 64, cudaMemcpyDeviceToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync((char*)pk->d_fixed + 192, b2->d_pts, 2 * 128, cudaMemcpyDeviceToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    zkb_bases_free(b1);
    zkb_bases_free(b2);
    if (e != cudaSuccess) throw cuda_fail(ctx, e, "fixed points");
    if ((rc = pk_upload_circuit(ctx, pk, cs, coeff_m))) throw rc;
    if (snark) {
      // coset shift of the witness map: the odd points of the size-2n domain are w_2n * <w_n>
      uint32_t w2n[8], onew[8];
      host_fr_root(d->log_n + 1, 0, w2n);
      hfr_words(HFr::one(), onew);
      cudaError_t e2 = cudaMalloc(&pk->d_shift, n * 32);
      if (e2 != cudaSuccess) throw cuda_fail(ctx, e2, "coset shift table");
      fr_powers_launch(pk->d_shift, w2n, onew, n, 1, ctx->stream);
      if ((e2 = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) throw cuda_fail(ctx, e2, "coset shift table");
    }
  } catch (int code) {
    zkb_pk_free(pk);
    return code;
  } catch (const std::bad_alloc&) {
    zkb_pk_free(pk);
    return set_error(ctx, ZKB_ENOMEM, "out of host memory");
  }
  *out = pk;
  return ZKB_OK;
}

// phase 1: witness -> the five MSM sums over this key's point ranges (pk->d_results) and h (pk->d_a)
static int prove_msm_phase(zkb_ctx* ctx, const zkb_pk* pk, const uint8_t* witness,
                           zkb_prove_timings* timings_opt, uint8_t* h_out_opt, uint8_t* results_out_opt) {
  const uint64_t n = pk->n, m = pk->m;
  cudaStream_t st = ctx->stream;
  enum { E_START, E_H2D, E_SPMV, E_NTT, E_A, E_B1, E_B2, E_L, E_H, E_COUNT };
  cudaEvent_t ev[E_COUNT];
  for (int i = 0; i < E_COUNT; i++) ZKB_CUDA(ctx, cudaEventCreate(&ev[i]));
  auto mark = [&](int i) { cudaEventRecord(ev[i], st); };
  int rc = ZKB_OK;
  cudaError_t e0 = cudaSuccess;
  uint32_t bad_flag = 0;
  do {
    mark(E_START);
    if ((e0 = ctx->msm.flag.ensure(16)) != cudaSuccess) { rc = cuda_fail(ctx, e0, "flag"); break; }
    cudaMemsetAsync(ctx->msm.flag.p, 0, 16, st);   // the MSM digit kernels flag a non-canonical witness
    cudaMemcpyAsync(pk->d_z, witness, m * 32, cudaMemcpyDefault, st);
    mark(E_H2D);
    const uint32_t bl_m = uint32_t((m + 255) / 256), bl_n = uint32_t((n + 255) / 256);
    fr_to_mont_copy_kernel<<<bl_m, 256, 0, st>>>((const Fr*)pk->d_z, (Fr*)pk->d_zm, m);
    count_launches(5);   // + three SpMV launches and the pointwise kernel below
    void* outs[3] = {pk->d_a, pk->d_b, pk->d_c};
    const bool snark = pk->h_basis == 1;
    for (int k = 0; k < (snark ? 2 : 3); k++)
      r1cs_spmv_kernel<<<bl_n, 256, 0, st>>>(pk->d_row_ptr[k], pk->d_col[k], pk->d_cidx[k],
                                             (const Fr*)pk->d_coeff_table, (const Fr*)pk->d_zm,
                                             (Fr*)outs[k], pk->ncons, n, k == 0 && !snark, pk->n_inst, pk->log_n, 0, 0, 0);
    if (snark) qap_product_kernel<<<bl_n, 256, 0, st>>>((const Fr*)pk->d_a, (const Fr*)pk->d_b, (Fr*)pk->d_c, n);
    mark(E_SPMV);
    if (snark) {
      // snarkjs witness map: a, b, c to coefficients, out to the odd points of the size-2n domain
      // (coset shift w_2n), a*b - c there — no division by Z, no way back: six transforms
      for (int k = 0; k < 3 && !rc; k++) rc = ntt_run(ctx, outs[k], outs[k], pk->log_n, 1, nullptr, nullptr, 0, 1, 1);
      for (int k = 0; k < 3 && !rc; k++) rc = ntt_run(ctx, outs[k], outs[k], pk->log_n, 0, pk->d_shift, nullptr, 0, 1, 1);
      if (rc) break;
      qap_pointwise_canonical_kernel<<<bl_n, 256, 0, st>>>((Fr*)pk->d_a, (const Fr*)pk->d_b, (const Fr*)pk->d_c, n);
      count_launches(1);
    } else {
      CosetTables* ct;
      if ((rc = coset_tables(ctx, pk->log_n, &ct))) break;
      for (int k = 0; k < 3 && !rc; k++) rc = ntt_run(ctx, outs[k], outs[k], pk->log_n, 1, nullptr, nullptr, 0, 1, 1);
      for (int k = 0; k < 3 && !rc; k++) rc = ntt_run(ctx, outs[k], outs[k], pk->log_n, 0, ct->fwd.p, nullptr, 0, 1, 1);
      if (rc) break;
      qap_pointwise_kernel<<<bl_n, 256, 0, st>>>((Fr*)pk->d_a, (const Fr*)pk->d_b, (const Fr*)pk->d_c, n);
      // coset iNTT with g^{-i} / Z(g) folded into the output table; result canonical = h
      if ((rc = ntt_run(ctx, pk->d_a, pk->d_a, pk->log_n, 1, nullptr, ct->inv_z.p, 0, 1, 0))) break;
    }
    mark(E_NTT);
    {
      // A, B1, L (G1) and B2 (G2) all take the witness as scalars: one digits + sort, four walks; the
      // H walk follows and ONE tail reduces the four G1 bucket sets (msm_exec_prove)
      const char* z = (const char*)pk->d_z;
      cudaEvent_t evs[4] = {ev[E_A], ev[E_B1], ev[E_L], ev[E_B2]};
      if ((rc = msm_exec_prove(ctx, pk->a_query, pk->b_g1_query, pk->l_query, pk->b_g2_query, pk->h_query,
                               z + 32 * pk->qfirst[0], pk->qcount[0], (const char*)pk->d_a + 32 * pk->qfirst[4],
                               pk->qcount[4], pk->d_results, evs))) break;
    }
    mark(E_H);
    if (results_out_opt)
      cudaMemcpyAsync(results_out_opt, pk->d_results, ZKB_INTERMEDIATE_BYTES, cudaMemcpyDeviceToHost, st);
    if (h_out_opt) cudaMemcpyAsync(h_out_opt, pk->d_a, (n - 1) * 32, cudaMemcpyDeviceToHost, st);   // snarkjs keys: the first n - 1 H scalars
    cudaMemcpyAsync(&bad_flag, ctx->msm.flag.p, 4, cudaMemcpyDeviceToHost, st);
    cudaError_t e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) {
      rc = cuda_fail(ctx, e, "groth16 prove");
      break;
    }
    if (bad_flag & 1) {
      cudaMemsetAsync(ctx->msm.flag.p, 0, 16, st);
      rc = set_error(ctx, ZKB_EINVAL, "non-canonical Fr element in the witness");
      break;
    }
    if (timings_opt) {
      // stream order: START H2D SPMV NTT | A B1 L B2 (shared sort, charged to A) | H
      const int order[9] = {E_START, E_H2D, E_SPMV, E_NTT, E_A, E_B1, E_L, E_B2, E_H};
      float* g[] = {&timings_opt->ms_h2d, &timings_opt->ms_spmv, &timings_opt->ms_ntt, &timings_opt->ms_msm_a,
                    &timings_opt->ms_msm_b1, &timings_opt->ms_msm_l, &timings_opt->ms_msm_b2, &timings_opt->ms_msm_h};
      for (int i = 0; i < 8; i++) cudaEventElapsedTime(g[i], ev[order[i]], ev[order[i + 1]]);
      cudaEventElapsedTime(&timings_opt->ms_total, ev[E_START], ev[E_H]);
    }
  } while (0);
  for (int i = 0; i < E_COUNT; i++) cudaEventDestroy(ev[i]);
  return rc;
}

// phase 2: MSM sums in pk->d_results + (r, s) -> proof
static int prove_assemble_phase(zkb_ctx* ctx, const zkb_pk* pk, const uint8_t r[32], const uint8_t s[32],
                                uint8_t* proof_out, float* ms_opt) {
  uint64_t w[4];
  std::memcpy(w, r, 32);
  bool ok = HFr::lt_mod(w);
  std::memcpy(w, s, 32);
  ok = ok && HFr::lt_mod(w);
  if (!ok) return set_error(ctx, ZKB_EINVAL, "non-canonical r / s");
  cudaStream_t st = ctx->stream;
  if (!ctx->assemble_on_device) {
    // the five sums come home (384 bytes) and the ladders run on host cores: see host_groth16_assemble
    uint8_t sums[ZKB_INTERMEDIATE_BYTES];
    auto t0 = std::chrono::steady_clock::now();
    ZKB_CUDA(ctx, cudaMemcpyAsync(sums, pk->d_results, ZKB_INTERMEDIATE_BYTES, cudaMemcpyDeviceToHost, st));
    cudaError_t e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(ctx, e, "groth16 msm phase");
    auto t1 = std::chrono::steady_clock::now();
    if (!pk->h_tables) pk->h_tables = host_assemble_tables_create(pk->h_fixed);   // null on failure: plain ladders
    if (!host_groth16_assemble(pk->h_fixed, sums, r, s, proof_out, pk->h_tables))
      return set_error(ctx, ZKB_EINVAL, "malformed MSM sums / key constants in the proof assembly");
    if (ms_opt) *ms_opt = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t1).count();
    (void)t0;
    return ZKB_OK;
  }
  cudaEvent_t e0, e1;
  ZKB_CUDA(ctx, cudaEventCreate(&e0));
  ZKB_CUDA(ctx, cudaEventCreate(&e1));
  cudaEventRecord(e0, st);
  cudaMemcpyAsync((char*)pk->d_proof + ZKB_PROOF_BYTES, r, 32, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync((char*)pk->d_proof + ZKB_PROOF_BYTES + 32, s, 32, cudaMemcpyHostToDevice, st);
  groth16_assemble_kernel<<<1, 128, 0, st>>>((const uint8_t*)pk->d_results, (const uint8_t*)pk->d_fixed,
                                             (const uint32_t*)((char*)pk->d_proof + ZKB_PROOF_BYTES),
                                             (uint8_t*)pk->d_proof);
  count_launches(1);
  cudaMemcpyAsync(proof_out, pk->d_proof, ZKB_PROOF_BYTES, cudaMemcpyDeviceToHost, st);
  cudaEventRecord(e1, st);
  cudaError_t e = cudaStreamSynchronize(st);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e == cudaSuccess && ms_opt) cudaEventElapsedTime(ms_opt, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "groth16 assemble");
  return ZKB_OK;
}

// witness -> a, b, c evaluation vectors (Montgomery, natural order) in pk->d_a / d_b / d_c
// (witness == nullptr: the caller has already staged the canonical witness in pk->d_z)
static int prove_witness_phase(zkb_ctx* ctx, const zkb_pk* pk, const uint8_t* witness,
                               uint32_t low_bits = 0, uint32_t world = 1, uint32_t rank = 0,
                               void* const* d_out_opt = nullptr) {
  const uint64_t n = pk->n, m = pk->m;
  cudaStream_t st = ctx->stream;
  if (witness) {
    ZKB_CUDA(ctx, ctx->msm.flag.ensure(16));
    ZKB_CUDA(ctx, cudaMemsetAsync(ctx->msm.flag.p, 0, 16, st));
    ZKB_CUDA(ctx, cudaMemcpyAsync(pk->d_z, witness, m * 32, cudaMemcpyDefault, st));
  }
  const uint32_t bl_m = uint32_t((m + 255) / 256);
  fr_to_mont_copy_kernel<<<bl_m, 256, 0, st>>>((const Fr*)pk->d_z, (Fr*)pk->d_zm, m);
  count_launches(4);   // + the three SpMV launches below
  void* outs[3] = {pk->d_a, pk->d_b, pk->d_c};
  uint64_t count = n;
  uint32_t lows = 0;
  if (low_bits) {
    lows = (1u << low_bits) / world;
    count = uint64_t(lows) << (pk->log_n - low_bits);
    for (int k = 0; k < 3; k++) outs[k] = d_out_opt[k];
  }
  const uint32_t bl = uint32_t((count + 255) / 256);
  for (int k = 0; k < 3; k++)
    r1cs_spmv_kernel<<<bl, 256, 0, st>>>(pk->d_row_ptr[k], pk->d_col[k], pk->d_cidx[k],
                                         (const Fr*)pk->d_coeff_table, (const Fr*)pk->d_zm,
                                         (Fr*)outs[k], pk->ncons, count, k == 0, pk->n_inst, pk->log_n,
                                         low_bits, lows, rank * lows);
  ZKB_CUDA(ctx, cudaGetLastError());
  return ZKB_OK;
}

// the five MSMs over pk's ranges into pk->d_results (device, asynchronous); h from d_h (canonical,
// pk's local h order)
static int prove_msm_dev_phase(zkb_ctx* ctx, const zkb_pk* pk, const void* d_h, cudaEvent_t* ev4_opt = nullptr) {
  const char* z = (const char*)pk->d_z;
  return msm_exec_prove(ctx, pk->a_query, pk->b_g1_query, pk->l_query, pk->b_g2_query, pk->h_query,
                        z + 32 * pk->qfirst[0], pk->qcount[0], d_h, pk->qcount[4], pk->d_results, ev4_opt);
}

// witness MSMs over pk's ranges + H MSM with h taken from d_h (canonical, pk's local h order)
static int prove_msm_only_phase(zkb_ctx* ctx, const zkb_pk* pk, const void* d_h, uint8_t* results_out) {
  int rc = prove_msm_dev_phase(ctx, pk, d_h);
  if (rc) return rc;
  uint32_t bad_flag = 0;
  ZKB_CUDA(ctx, cudaMemcpyAsync(results_out, pk->d_results, ZKB_INTERMEDIATE_BYTES, cudaMemcpyDeviceToHost, ctx->stream));
  ZKB_CUDA(ctx, cudaMemcpyAsync(&bad_flag, ctx->msm.flag.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
  ZKB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ZKB_CUDA(ctx, cudaGetLastError());
  if (bad_flag & 1) {
    cudaMemsetAsync(ctx->msm.flag.p, 0, 16, ctx->stream);
    return set_error(ctx, ZKB_EINVAL, "non-canonical Fr element in the witness");
  }
  return ZKB_OK;
}

#define ZKB_PK_ENTER(ctx, pk)                                                             \
  if (!(ctx)) return ZKB_EINVAL;                                                          \
  BusyGuard guard(ctx);                                                                   \
  if (!guard.ok) return ZKB_EBUSY;                                                        \
  if (!(pk)) return set_error(ctx, ZKB_EINVAL, "null proving key");                       \
  if ((pk)->device != (ctx)->device)                                                      \
    return set_error(ctx, ZKB_EINVAL, "proving key lives on another device");            \
  PkGuard pk_guard(pk);                                                                   \
  if (!pk_guard.ok) return set_error(ctx, ZKB_EBUSY, "proving key is in use by another prove call"); \
  ZKB_CUDA(ctx, cudaSetDevice((ctx)->device));

int zkb_groth16_prove(zkb_ctx* ctx, const zkb_pk* pk, const uint8_t* witness, const uint8_t r[32],
                      const uint8_t s[32], uint8_t proof_out[ZKB_PROOF_BYTES],
                      uint8_t* intermediates_opt, uint8_t* h_out_opt,
                      zkb_prove_timings* timings_opt) {
  ZKB_PK_ENTER(ctx, pk);
  if (!witness || !r || !s || !proof_out) return set_error(ctx, ZKB_EINVAL, "null argument");
  if (pk->world != 1)
    return set_error(ctx, ZKB_EINVAL, "sharded key: use zkb_groth16_prove_partial + zkb_groth16_assemble");
  if (pk->h_low_bits)
    return set_error(ctx, ZKB_EINVAL, "key stores h_query in the four-step layout: use the distributed prover");
  int rc = prove_msm_phase(ctx, pk, witness, timings_opt, h_out_opt, intermediates_opt);
  if (rc) {
    std::memset(proof_out, 0, ZKB_PROOF_BYTES);
    return rc;
  }
  float ms = 0;
  rc = prove_assemble_phase(ctx, pk, r, s, proof_out, &ms);
  if (!rc && timings_opt) timings_opt->ms_total += ms;
  return rc;
}

int zkb_groth16_prove_partial(zkb_ctx* ctx, const zkb_pk* pk, const uint8_t* witness,
                              uint8_t partial_out[ZKB_INTERMEDIATE_BYTES], zkb_prove_timings* timings_opt) {
  ZKB_PK_ENTER(ctx, pk);
  if (!witness || !partial_out) return set_error(ctx, ZKB_EINVAL, "null argument");
  if (pk->h_low_bits)   // the single-GPU witness map leaves h in natural order: it would meet the wrong bases
    return set_error(ctx, ZKB_EINVAL, "key stores h_query in the four-step layout: use the distributed prover");
  return prove_msm_phase(ctx, pk, witness, timings_opt, nullptr, partial_out);
}

int zkb_groth16_witness_dev(zkb_ctx* ctx, const zkb_pk* pk, const uint8_t* witness, void* d_abc_out[3]) {
  ZKB_PK_ENTER(ctx, pk);
  if (!witness || !d_abc_out) return set_error(ctx, ZKB_EINVAL, "null argument");
  if (pk->h_basis) return set_error(ctx, ZKB_EINVAL, "snarkjs-basis key: zkb_groth16_prove / _prove_partial only");
  int rc = prove_witness_phase(ctx, pk, witness);
  d_abc_out[0] = pk->d_a;
  d_abc_out[1] = pk->d_b;
  d_abc_out[2] = pk->d_c;
  return rc;
}

int zkb_groth16_witness_layout_dev(zkb_ctx* ctx, const zkb_pk* pk, const uint8_t* witness, uint32_t low_bits,
                                   uint32_t world, uint32_t rank, void* const d_abc_local[3]) {
  ZKB_PK_ENTER(ctx, pk);
  if (!witness || !d_abc_local || !d_abc_local[0] || !d_abc_local[1] || !d_abc_local[2])
    return set_error(ctx, ZKB_EINVAL, "null argument");
  if (low_bits == 0 || low_bits >= pk->log_n || world == 0 || rank >= world || ((1u << low_bits) % world))
    return set_error(ctx, ZKB_EINVAL, "bad layout argument");
  if (pk->h_basis) return set_error(ctx, ZKB_EINVAL, "snarkjs-basis key: zkb_groth16_prove / _prove_partial only");
  return prove_witness_phase(ctx, pk, witness, low_bits, world, rank, d_abc_local);
}

int zkb_groth16_msm_dev(zkb_ctx* ctx, const zkb_pk* pk, const void* d_h,
                        uint8_t partial_out[ZKB_INTERMEDIATE_BYTES]) {
  ZKB_PK_ENTER(ctx, pk);
  if (!d_h || !partial_out) return set_error(ctx, ZKB_EINVAL, "null argument");
  return prove_msm_only_phase(ctx, pk, d_h, partial_out);
}

// Layout of the evaluation vectors in the distributed witness map.  Every transform of the chain maps
// layout d0 -> L - d0 -> d0 ..., i.e. it is one batch of size-2^d0 transforms plus one batch of
// size-2^(L-d0) transforms.  d0 = 8 makes the first kind a SINGLE pass of the NTT kernel, so a
// distributed transform costs as many passes as the single-GPU one (2^22: 1 + 2 instead of the
// 2 + 2 of the balanced split 11 + 11); small circuits keep the balanced split.
static uint32_t dist_low_bits(uint32_t log_n) { return log_n / 2 > 8 ? 8 : log_n / 2; }

// ---- the whole proof over the ranks of ctx's communicator (collective; BASELINE.json configs[4])
int zkb_pk_synthetic_dist(zkb_ctx* ctx, uint32_t log_n, uint64_t seed, zkb_pk** out, uint8_t vk_out[ZKB_VK_BYTES]) {
  if (!ctx) return ZKB_EINVAL;
  if (!ctx->comm) return set_error(ctx, ZKB_EINVAL, "no communicator (zkb_comm_init first)");
  const uint32_t world = uint32_t(ctx->comm->world), d0 = dist_low_bits(log_n);
  if (log_n < 4 || ((1u << d0) % world))
    return set_error(ctx, ZKB_EINVAL, "circuit too small for this many ranks (need 2^(log_n/2) >= world)");
  return zkb_pk_synthetic_shard_ex(ctx, log_n, seed, uint32_t(ctx->comm->rank), world, log_n - d0, out, vk_out);
}

int zkb_groth16_prove_dist(zkb_ctx* ctx, const zkb_pk* pk, const uint8_t* witness, const uint8_t r[32],
                           const uint8_t s[32], uint8_t proof_out[ZKB_PROOF_BYTES], zkb_prove_timings* timings_opt) {
  ZKB_PK_ENTER(ctx, pk);
  if (!witness || !r || !s || !proof_out) return set_error(ctx, ZKB_EINVAL, "null argument");
  Comm* cm = ctx->comm;
  if (!cm) return set_error(ctx, ZKB_EINVAL, "no communicator (zkb_comm_init first)");
  const uint32_t L = pk->log_n, d0 = dist_low_bits(L), d1 = L - d0, G = uint32_t(cm->world), g = uint32_t(cm->rank);
  if (pk->world != G || pk->rank != g || pk->h_low_bits != d1)
    return set_error(ctx, ZKB_EINVAL, "key was not built for this communicator (use zkb_pk_synthetic_dist)");
  if (pk->h_basis) return set_error(ctx, ZKB_EINVAL, "snarkjs-basis key: single-GPU prove only");
  const uint64_t n = pk->n, m = pk->m, n_loc = n / G;
  cudaStream_t st = ctx->stream;
  enum { E_START, E_WIT, E_SPMV, E_NTT, E_A, E_B1, E_L, E_B2, E_MSMH, E_END, E_COUNT };
  cudaEvent_t ev[E_COUNT];
  for (int i = 0; i < E_COUNT; i++) ZKB_CUDA(ctx, cudaEventCreate(&ev[i]));
  int rc = ZKB_OK;
  do {
    cudaEventRecord(ev[E_START], st);
    cudaError_t e0 = ctx->msm.flag.ensure(16);
    if (e0 == cudaSuccess) e0 = ctx->dist_small.ensure(size_t(kMaxRanks) * 512 + 1024);
    if (e0 != cudaSuccess) { rc = cuda_fail(ctx, e0, "scratch"); break; }
    cudaMemsetAsync(ctx->msm.flag.p, 0, 16, st);
    // 1. witness: every rank uploads 1/G of it over PCIe, the slices travel over NVLink
    cudaPointerAttributes attr;
    const bool on_device = cudaPointerGetAttributes(&attr, witness) == cudaSuccess && attr.type == cudaMemoryTypeDevice;
    cudaGetLastError();
    if (on_device || G == 1) {
      cudaMemcpyAsync(pk->d_z, witness, m * 32, cudaMemcpyDefault, st);
    } else {
      const uint64_t per = m * 32 / G;
      char* mine = (char*)pk->d_z + g * per;
      cudaMemcpyAsync(mine, witness + g * per, per, cudaMemcpyHostToDevice, st);
      if ((rc = comm_allgather(ctx, mine, pk->d_z, per))) break;
    }
    cudaEventRecord(ev[E_WIT], st);
    // 2. A z, B z, C z: only the rows of this rank's layout-d0 block
    void* bufs[3] = {pk->d_a, pk->d_b, pk->d_c};
    if ((rc = prove_witness_phase(ctx, pk, nullptr, d0, G, g, bufs))) break;
    cudaEventRecord(ev[E_SPMV], st);
    // 3. the witness map: 7 four-step transforms, coset shifts and 1/Z(g) folded into index tables
    const HFr gen = HFr::from_u64(5);
    const HFr zinv = (hpow_u64(gen, n) - HFr::one()).inverse();
    const void *t_coset = nullptr, *t_uncoset = nullptr;
    if ((rc = dist_index_power_table(ctx, L, d1, gen, HFr::one(), &t_coset))) break;
    if ((rc = dist_index_power_table(ctx, L, d1, gen.inverse(), zinv, &t_uncoset))) break;
    for (int k = 0; k < 3 && !rc; k++) rc = dist_ntt_transform(ctx, bufs[k], L, d0, 1, nullptr, t_coset);
    for (int k = 0; k < 3 && !rc; k++) rc = dist_ntt_transform(ctx, bufs[k], L, d1, 0, nullptr, nullptr);
    if (rc) break;
    fr_mul_sub_launch(pk->d_a, pk->d_b, pk->d_c, n_loc, st);
    if ((rc = dist_ntt_transform(ctx, pk->d_a, L, d0, 1, nullptr, t_uncoset))) break;
    fr_convert_launch(pk->d_a, n_loc, 0, st);
    cudaEventRecord(ev[E_NTT], st);
    // 4. the five MSMs over this rank's point ranges
    if ((rc = prove_msm_dev_phase(ctx, pk, pk->d_a, &ev[E_A]))) break;
    cudaEventRecord(ev[E_MSMH], st);
    // 5. 384 bytes per rank: all-gather, slot-wise sums, assemble — all on the device
    static const uint32_t offs[5] = {0, 64, 128, 256, 320}, g2s[5] = {0, 0, 1, 0, 0};
    if ((rc = dist_sum_blocks(ctx, pk->d_results, ZKB_INTERMEDIATE_BYTES, offs, g2s, 5, pk->d_results))) break;
    // a non-canonical witness element is seen only by the rank whose MSM shard holds it: share the flags
    char* fl = (char*)ctx->dist_small.p + 512;
    cudaMemcpyAsync(fl, ctx->msm.flag.p, 4, cudaMemcpyDeviceToDevice, st);
    if ((rc = comm_allgather(ctx, fl, fl + 64, 4))) break;
    uint32_t flags[kMaxRanks] = {};
    cudaMemcpyAsync(flags, fl + 64, 4 * G, cudaMemcpyDeviceToHost, st);
    float ms_asm = 0;
    if ((rc = prove_assemble_phase(ctx, pk, r, s, proof_out, &ms_asm))) break;   // synchronises
    cudaEventRecord(ev[E_END], st);
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { rc = cuda_fail(ctx, e, "distributed prove"); break; }
    uint32_t bad = 0;
    for (uint32_t i = 0; i < G; i++) bad |= flags[i];
    if (bad & 1) {
      cudaMemsetAsync(ctx->msm.flag.p, 0, 16, st);
      std::memset(proof_out, 0, ZKB_PROOF_BYTES);
      rc = set_error(ctx, ZKB_EINVAL, "non-canonical Fr element in the witness");
      break;
    }
    if (timings_opt) {
      std::memset(timings_opt, 0, sizeof(*timings_opt));
      cudaEventElapsedTime(&timings_opt->ms_h2d, ev[E_START], ev[E_WIT]);
      cudaEventElapsedTime(&timings_opt->ms_spmv, ev[E_WIT], ev[E_SPMV]);
      cudaEventElapsedTime(&timings_opt->ms_ntt, ev[E_SPMV], ev[E_NTT]);
      cudaEventElapsedTime(&timings_opt->ms_msm_a, ev[E_NTT], ev[E_A]);        // shared sort + walk
      cudaEventElapsedTime(&timings_opt->ms_msm_b1, ev[E_A], ev[E_B1]);
      cudaEventElapsedTime(&timings_opt->ms_msm_l, ev[E_B1], ev[E_L]);
      cudaEventElapsedTime(&timings_opt->ms_msm_b2, ev[E_L], ev[E_B2]);
      cudaEventElapsedTime(&timings_opt->ms_msm_h, ev[E_B2], ev[E_MSMH]);      // sort + walk of H, then both tails
      cudaEventElapsedTime(&timings_opt->ms_total, ev[E_START], ev[E_END]);
    }
  } while (0);
  for (int i = 0; i < E_COUNT; i++) cudaEventDestroy(ev[i]);
  return rc;
}

int zkb_groth16_assemble(zkb_ctx* ctx, const zkb_pk* pk, const uint8_t sums[ZKB_INTERMEDIATE_BYTES],
                         const uint8_t r[32], const uint8_t s[32], uint8_t proof_out[ZKB_PROOF_BYTES]) {
  ZKB_PK_ENTER(ctx, pk);
  if (!sums || !r || !s || !proof_out) return set_error(ctx, ZKB_EINVAL, "null argument");
  ZKB_CUDA(ctx, cudaMemcpyAsync(pk->d_results, sums, ZKB_INTERMEDIATE_BYTES, cudaMemcpyHostToDevice, ctx->stream));
  return prove_assemble_phase(ctx, pk, r, s, proof_out, nullptr);
}

}  // extern "C"
