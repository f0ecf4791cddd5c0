/* zkb.h — C ABI of libzkb.so, the B200-native BN254 Groth16 proving backend.
 *
 * This is the drop-in boundary UNDER eigen-zeth's prover service: the reference reaches its
 * prover only over gRPC (reference proto/prover/v1/prover.proto:9-11, client at
 * src/prover/provider.rs:603-705) and has no FFI of its own, so each entry point
 * below cites the piece of the reference-side contract it serves rather than a function it
 * replaces.  A Rust `-sys` crate binds these symbols 1:1 through build.rs.
 *
 * Conventions
 *  - every function returns 0 (ZKB_OK) or a negative error code; zkb_last_error(ctx) explains;
 *  - no exception or abort crosses this boundary;
 *  - field elements: 32 bytes, little-endian, canonical (< modulus), NOT Montgomery;
 *  - G1 affine: x||y (64 B); G2 affine: x.c0||x.c1||y.c0||y.c1 (128 B); all-zero = infinity.
 *    (c0 first is the order of reference proof/proof.json "pi_b", read unswapped by
 *     src/settlement/ethereum/mod.rs:451-466.)
 *  - a zkb_ctx is bound to ONE GPU and is not re-entrant: one call at a time per ctx (ZKB_EBUSY
 *    otherwise).  Several GPUs = several contexts (one per process, or one per host thread), joined
 *    into a communicator with zkb_comm_init; the zkb_*_dist calls are then collective;
 *  - a proving key is used by one prove call at a time (its work vectors live in the key): a second
 *    concurrent call on the same key returns ZKB_EBUSY;
 *  - uploaded bases / proving-key queries must be on the curve (checked on the GPU, ZKB_EINVAL
 *    otherwise); G2 points are NOT checked for membership of the order-r subgroup (a ceremony key
 *    is expected to have been validated by its own tooling);
 *  - `*_dev` variants take raw device pointers on the ctx's device and enqueue on the ctx's
 *    stream (zkb_ctx_stream) without synchronising, so callers can time them with CUDA events.
 */
#ifndef ZKB_H
#define ZKB_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#pragma GCC visibility push(default)

#define ZKB_OK 0
#define ZKB_EINVAL (-1)   /* bad argument (size, non-canonical field element, null pointer) */
#define ZKB_ECUDA (-2)    /* CUDA runtime error (message in zkb_last_error) */
#define ZKB_ENOMEM (-3)
#define ZKB_EBUSY (-4)    /* ctx already executing a call on another thread */
#define ZKB_ENOGPU (-5)   /* no usable sm_100 device: there is no CPU fallback */
#define ZKB_EVERIFY (-6)  /* pairing check failed */
#define ZKB_ECOMM (-7)    /* NCCL / peer-memory failure in a multi-GPU call (message in zkb_last_error) */

typedef struct zkb_ctx zkb_ctx;
typedef struct zkb_bases zkb_bases;
typedef struct zkb_pk zkb_pk;

/* ---- context ------------------------------------------------------------------------------- */
int zkb_ctx_create(int device, zkb_ctx** out);
void zkb_ctx_destroy(zkb_ctx* ctx);
const char* zkb_last_error(const zkb_ctx* ctx);
void* zkb_ctx_stream(zkb_ctx* ctx);                 /* cudaStream_t */
int zkb_ctx_sync(zkb_ctx* ctx);
/* tunables: "msm_window" (0 = auto), "ntt_full_twiddles" (0/1), "msm_profile" (0/1),
 * "msm_host_chunks" (0 = auto, 1..16), "pk_precompute" (0/1: window multiples for proving keys),
 * "pk_window" (0 = by size; 8..24: window width of the queries of keys created afterwards),
 * "assemble_on_device" (0/1, see zkb_groth16_assemble_host),
 * "msm_sort" (process-wide, default 0 = the library's own bucket sort fused with digit extraction;
 * 1 = digits kernel + cub::DeviceRadixSort, kept as the comparison baseline),
 * "msm_g2_pair" (process-wide bit mask, default 2: bit 1 = the latency-bound G2 kernels around the
 * bucket walk run one Fq2 operation per lane PAIR; bit 0 = the walk itself does too — measured 5 %
 * slower than one thread per point on B200, kept for comparison) */
int zkb_ctx_set(zkb_ctx* ctx, const char* key, int64_t value);
const char* zkb_version(void);
/* page-locked (pinned, portable) host memory for witnesses / scalars: H2D copies from it run at PCIe
 * speed and overlap with compute; plain malloc'ed buffers are accepted everywhere too (slower) */
int zkb_host_alloc(uint64_t bytes, void** out);
void zkb_host_free(void* p);

/* ---- Fr NTT  (serves the 7 transforms of GenFinalProof) ---------------- */
/* in place on a host buffer of 2^log_n canonical elements; ark-poly Radix2EvaluationDomain
 * semantics: natural order, inverse carries 1/n, coset uses the shift g = 5. */
int zkb_fr_ntt(zkb_ctx* ctx, uint8_t* data, uint32_t log_n, int inverse, int coset);
/* device-resident variant: d_in -> d_out (may alias), elements canonical unless montgomery != 0 */
int zkb_fr_ntt_dev(zkb_ctx* ctx, const void* d_in, void* d_out, uint32_t log_n, int inverse,
                   int coset, int montgomery);

/* Batched / table building blocks (the multi-GPU four-step NTT itself is zkb_fr_ntt_dist_dev below;
 * eigen-zeth_b200/dist_ntt.py composes the same steps above this layer for the emulated-rank
 * tests).  Device buffers, Montgomery form, asynchronous on the ctx stream.
 *  - `batch` contiguous transforms of 2^log_size elements; tables (optional, batch * 2^log_size
 *    entries, Montgomery) multiply the input / output element-wise; skip_ninv leaves 1/size out;
 *  - table mode 0: out[hi*cols+lo] = scale * base^((row0+hi)*lo); mode 1: scale * base^((row0+hi) + lo*mul)
 *    (base, scale canonical; output Montgomery);
 *  - mul_sub: a[i] = a[i]*b[i] - c[i];   convert: canonical <-> Montgomery in place. */
int zkb_fr_ntt_batch_dev(zkb_ctx* ctx, const void* d_in, void* d_out, uint32_t log_size, uint64_t batch,
                         int inverse, int skip_ninv, const void* d_pre_table, const void* d_post_table);
/* this rank's block of "layout low_bits" (see zkb_fr_ntt_dist_dev) out of a natural-order device vector */
int zkb_fr_gather_layout_dev(zkb_ctx* ctx, const void* d_full, void* d_out, uint32_t log_n, uint32_t low_bits,
                             uint32_t world, uint32_t rank);
int zkb_fr_table_dev(zkb_ctx* ctx, void* d_out, const uint8_t base[32], const uint8_t scale[32], uint64_t rows,
                     uint64_t cols, uint64_t row0, uint64_t mul, int mode);
int zkb_fr_mul_sub_dev(zkb_ctx* ctx, void* d_a, const void* d_b, const void* d_c, uint64_t n);
int zkb_fr_convert_dev(zkb_ctx* ctx, void* d_data, uint64_t n, int to_montgomery);
/* canonical 2^log_n-th root of unity of Fr (ark-ff TWO_ADIC_ROOT_OF_UNITY^(2^(28-log_n))) or its inverse */
void zkb_fr_root_of_unity(uint32_t log_n, int inverse, uint8_t out[32]);

/* ---- MSM  (serves the A/B1/L/H G1 MSMs and the B2 G2 MSM) ---------- */
int zkb_bases_upload_g1(zkb_ctx* ctx, const uint8_t* affine, uint64_t n, zkb_bases** out);
int zkb_bases_upload_g2(zkb_ctx* ctx, const uint8_t* affine, uint64_t n, zkb_bases** out);
/* deterministic synthetic bases P_i = k_i * G (k_i from xoshiro256**(seed)), generated on the GPU */
int zkb_bases_synthetic_g1(zkb_ctx* ctx, uint64_t n, uint64_t seed, zkb_bases** out);
int zkb_bases_synthetic_g2(zkb_ctx* ctx, uint64_t n, uint64_t seed, zkb_bases** out);
/* copy `count` bases starting at `first` back to the host in canonical form */
int zkb_bases_download(zkb_ctx* ctx, const zkb_bases* b, uint64_t first, uint64_t count,
                       uint8_t* out);
uint64_t zkb_bases_len(const zkb_bases* b);
/* Expand the bases in HBM to their window multiples 2^{c w} P_i (ceil(255/c) copies: the B200's
 * 180 GB is the resource being spent).  Every window then feeds ONE shared bucket set: no Horner
 * pass, a bucket reduction of a single window, and a wider window (fewer additions) for the same
 * bucket count.  window_bits = 0 picks the width from the number of points.  Proving-key queries
 * are fixed, so this is done once per key.  MSM results are unchanged (bit-exact). */
int zkb_bases_precompute(zkb_ctx* ctx, zkb_bases* b, int window_bits);
int zkb_bases_window(const zkb_bases* b);          /* 0 = not precomputed */
void zkb_bases_free(zkb_bases* b);
/* sum_i scalars[i] * bases[first + i], i < n; host scalars; result canonical affine (64 / 128 B) */
int zkb_msm_g1(zkb_ctx* ctx, const zkb_bases* b, uint64_t first, const uint8_t* scalars,
               uint64_t n, uint8_t out[64]);
int zkb_msm_g2(zkb_ctx* ctx, const zkb_bases* b, uint64_t first, const uint8_t* scalars,
               uint64_t n, uint8_t out[128]);
/* device-resident scalars (canonical); d_out is a device buffer of 64 / 128 bytes; asynchronous */
int zkb_msm_g1_dev(zkb_ctx* ctx, const zkb_bases* b, uint64_t first, const void* d_scalars,
                   uint64_t n, void* d_out);
int zkb_msm_g2_dev(zkb_ctx* ctx, const zkb_bases* b, uint64_t first, const void* d_scalars,
                   uint64_t n, void* d_out);
/* sum of n affine points (the per-rank partial results of a point-range-sharded MSM) */
int zkb_g1_sum(zkb_ctx* ctx, const uint8_t* points, uint32_t n, uint8_t out[64]);
int zkb_g2_sum(zkb_ctx* ctx, const uint8_t* points, uint32_t n, uint8_t out[128]);
/* the same on device buffers (e.g. the output of an NCCL all-gather), asynchronous */
int zkb_g1_sum_dev(zkb_ctx* ctx, const void* d_points, uint32_t n, void* d_out);
int zkb_g2_sum_dev(zkb_ctx* ctx, const void* d_points, uint32_t n, void* d_out);

/* ---- several GPUs  (BASELINE.json configs[3], configs[4]) ------------------------
 * The reference has no distributed backend: its operator knows ONE prover address
 * (src/config/env.rs:21, scripts/start-l2.sh:3).  Behind that address one proof can be spread over
 * the GPUs of a box: one zkb_ctx per GPU, joined into a communicator, and the zkb_*_dist calls are
 * COLLECTIVE — every rank calls them, in the same order, from its own host thread or process.
 *  - zkb_comm_init: one process (or thread) per rank; rendezvous through NCCL (libnccl.so.2 is
 *    loaded at this call, not before): rank 0 makes an id with zkb_comm_unique_id and hands it to
 *    the others by any means (file, socket, torch.distributed);
 *  - zkb_comm_init_local: all ranks inside ONE process (the prover service: one host thread per
 *    GPU); no NCCL at all — peer pointers and CUDA events.  Contexts may share a device (tests).
 * world must be a power of two <= 16.
 * The exchange of the four-step NTT is fused into the NTT kernels: the last pass of the first step
 * stores each element straight into the destination rank's receive window (peer memory over
 * NVLink; CUDA IPC between processes), the first pass of the second step reads the window
 * transposed.  zkb_comm_transport: 1 = those peer-memory windows, 2 = NCCL send/recv staging
 * (fallback when peer mapping is refused), 0 = no exchange has run yet. */
#define ZKB_COMM_ID_BYTES 128
int zkb_comm_unique_id(uint8_t id_out[ZKB_COMM_ID_BYTES]);
int zkb_comm_init(zkb_ctx* ctx, const uint8_t id[ZKB_COMM_ID_BYTES], int rank, int world);
int zkb_comm_init_local(zkb_ctx* const* ctxs, int world);
int zkb_comm_rank(const zkb_ctx* ctx);
int zkb_comm_world(const zkb_ctx* ctx);
int zkb_comm_transport(const zkb_ctx* ctx);
int zkb_comm_barrier(zkb_ctx* ctx);                 /* stream-ordered barrier over the ranks */
void zkb_comm_destroy(zkb_ctx* ctx);
/* One four-step transform of a length-2^log_n vector held in "layout low_bits": index = low +
 * 2^low_bits * high, rank g owns the lows [g * 2^low_bits / world, (g+1) * 2^low_bits / world) and
 * all highs, local array [low_loc][high] (2^log_n / world elements, Montgomery form, device).
 * In place; the result is in layout log_n - low_bits, natural NTT order (ark-poly semantics, an
 * inverse carries 1/n).  Tables (optional, device, Montgomery, 2^log_n / world entries) multiply
 * the input / output element-wise in the input / output layout. */
int zkb_fr_ntt_dist_dev(zkb_ctx* ctx, void* d_data, uint32_t log_n, uint32_t low_bits, int inverse,
                        const void* d_pre_table, const void* d_post_table);
/* table[pos] = scale * base^(global index of pos) for this rank's block of layout low_bits (coset
 * shifts); built once, cached and owned by the context */
int zkb_fr_index_power_table_dist(zkb_ctx* ctx, uint32_t log_n, uint32_t low_bits, const uint8_t base[32],
                                  const uint8_t scale[32], const void** d_table_out);
/* MSM sharded by point range: every rank passes ITS bases and scalars, every rank receives the
 * total.  The only exchange is an all-gather of one affine point per rank. */
int zkb_msm_g1_dist_dev(zkb_ctx* ctx, const zkb_bases* b, uint64_t first, const void* d_scalars,
                        uint64_t n, void* d_out);
int zkb_msm_g2_dist_dev(zkb_ctx* ctx, const zkb_bases* b, uint64_t first, const void* d_scalars,
                        uint64_t n, void* d_out);
int zkb_msm_g1_dist(zkb_ctx* ctx, const zkb_bases* b, uint64_t first, const uint8_t* scalars,
                    uint64_t n, uint8_t out[64]);
int zkb_msm_g2_dist(zkb_ctx* ctx, const zkb_bases* b, uint64_t first, const uint8_t* scalars,
                    uint64_t n, uint8_t out[128]);

/* ---- Groth16  (GenFinalProofRequest -> FinalProof; prover.proto:130-148) --------------------- */
/* verifying key layout: alpha_g1 64 | beta_g2 128 | gamma_g2 128 | delta_g2 128 | ic0 64 | ic1 64 */
#define ZKB_VK_BYTES 576
#define ZKB_PROOF_BYTES 256 /* A 64 | B 128 | C 64 */
/* synthetic chain circuit with 2^log_n - 2 constraints, 2^log_n variables, 1 public input;
 * toxic waste from xoshiro256**(seed); queries are computed on the GPU and stay resident. */
int zkb_pk_synthetic(zkb_ctx* ctx, uint32_t log_n, uint64_t seed, zkb_pk** out,
                     uint8_t vk_out[ZKB_VK_BYTES]);
/* the same key, but only the point range of rank `rank` out of `world` of every query stays on
 * this GPU (one process per GPU; the MSM partial sums are combined above this layer) */
int zkb_pk_synthetic_shard(zkb_ctx* ctx, uint32_t log_n, uint64_t seed, uint32_t rank, uint32_t world,
                           zkb_pk** out, uint8_t vk_out[ZKB_VK_BYTES]);
/* An externally generated key (f4 of SURVEY.md 8f: circuits other than the synthetic chain).
 * Matrices are CSR over the variables z = (1, public..., private...): row_ptr has n_constraints + 1
 * entries, coeff holds one canonical 32-byte Fr per non-zero.  Queries are the fields of
 * ark-groth16's ProvingKey as canonical affine bytes (all-zero = infinity): a_query, b_g1_query,
 * b_g2_query have n_vars points, l_query n_vars - n_public - 1, h_query 2^log_n - 1 (h_basis 0;
 * see h_basis for snarkjs keys).  The domain
 * must satisfy n_constraints + n_public + 1 <= 2^log_n (ark-groth16 appends one input-consistency
 * row per instance variable).  Everything is copied; the caller's buffers can be freed after. */
typedef struct zkb_pk_desc {
  uint32_t log_n;
  uint32_t n_public;
  uint64_t n_vars;
  uint64_t n_constraints;
  const uint32_t* row_ptr[3];   /* A, B, C */
  const uint32_t* col[3];
  const uint8_t* coeff[3];
  const uint8_t *a_query, *b_g1_query, *b_g2_query, *l_query, *h_query;
  const uint8_t *alpha_g1, *beta_g1, *delta_g1;   /* 64 B each */
  const uint8_t *beta_g2, *delta_g2;              /* 128 B each */
  /* 0: ark-groth16 (above).  1: snarkjs (.zkey): h_query holds 2^log_n points L_{2i+1}(tau) / delta
   * (Lagrange basis of the size-2^(log_n+1) domain at its odd points) and the prover evaluates
   * a * b - c on those points instead of dividing by the vanishing polynomial; matrix A already
   * contains the input rows (n_constraints counts them, n_constraints <= 2^log_n) and matrix C may be
   * null (row_ptr[2] == NULL: c = a * b on the domain).  Same proof bytes as h_basis 0 for the same
   * circuit, setup, witness and (r, s).  Single-GPU prove only. */
  uint32_t h_basis;
} zkb_pk_desc;
int zkb_pk_load(zkb_ctx* ctx, const zkb_pk_desc* desc, zkb_pk** out);
/* as above, with h_query held in the output layout of the four-step NTT: local [low_loc][high]
 * is the base of coefficient (rank * 2^h_low_bits / world + low_loc) + (high << h_low_bits) */
int zkb_pk_synthetic_shard_ex(zkb_ctx* ctx, uint32_t log_n, uint64_t seed, uint32_t rank, uint32_t world,
                              uint32_t h_low_bits, zkb_pk** out, uint8_t vk_out[ZKB_VK_BYTES]);
void zkb_pk_free(zkb_pk* pk);
/* borrowed handle to one query of the key: 0 a_query, 1 b_g1_query, 2 b_g2_query, 3 l_query,
 * 4 h_query (ark-groth16 ProvingKey field names); valid until zkb_pk_free; do not free it.
 * l_query is stored at the full witness length: its first 2 entries (the instance variables, which
 * ark-groth16 leaves out) are the point at infinity, so all witness MSMs share one scalar sort. */
const zkb_bases* zkb_pk_query(const zkb_pk* pk, int which);
/* number of variables m (the witness passed to the prove calls must hold exactly 32 * m bytes) and
 * log2 of the QAP domain */
uint64_t zkb_pk_n_vars(const zkb_pk* pk);
uint32_t zkb_pk_log_n(const zkb_pk* pk);
/* full witness assignment z (2^log_n canonical Fr, z[0] = 1, z[1] = public input) from two seeds */
int zkb_synthetic_witness(uint32_t log_n, const uint8_t w0[32], const uint8_t w1[32],
                          uint8_t* z_out);
/* device times: ms_msm_a = the witness digits + sort (shared by a, b_g1, l, b_g2) and the a_query
 * walk; ms_msm_b1 / _l / _b2 = the bucket walks of those queries; ms_msm_h = sort + walk of the
 * h_query MSM, then the bucket reductions (one shared by the four G1 MSMs, one for G2) */
typedef struct zkb_prove_timings {
  float ms_h2d, ms_spmv, ms_ntt, ms_msm_a, ms_msm_b1, ms_msm_b2, ms_msm_l, ms_msm_h, ms_total;
} zkb_prove_timings;
/* intermediates (optional, may be NULL): msm_a 64 | msm_b1 64 | msm_b2 128 | msm_l 64 | msm_h 64 */
#define ZKB_INTERMEDIATE_BYTES 384
int zkb_groth16_prove(zkb_ctx* ctx, const zkb_pk* pk, const uint8_t* witness, const uint8_t r[32],
                      const uint8_t s[32], uint8_t proof_out[ZKB_PROOF_BYTES],
                      uint8_t* intermediates_opt, uint8_t* h_out_opt,
                      zkb_prove_timings* timings_opt);
/* multi-GPU composition: every rank computes the five MSM sums over ITS point ranges
 * (msm_a 64 | msm_b1 64 | msm_b2 128 | msm_l 64 | msm_h 64); the ranks' partials are added slot by
 * slot (zkb_g1_sum / zkb_g2_sum after an all-gather) and any rank assembles the proof */
int zkb_groth16_prove_partial(zkb_ctx* ctx, const zkb_pk* pk, const uint8_t* witness,
                              uint8_t partial_out[ZKB_INTERMEDIATE_BYTES], zkb_prove_timings* timings_opt);
/* the prover split for the distributed witness map: (1) witness -> A z, B z, C z evaluation vectors
 * (device, Montgomery, natural order; pointers returned in d_abc_out, owned by the key);
 * (2) ... the caller runs the NTT chain (dist_ntt.py) ...; (3) the five MSMs with h read from d_h
 * (canonical, in the key's local h order) -> this rank's 384-byte partial sums */
/* `witness` may be a host pointer or a device pointer on the ctx's GPU (e.g. the output of an
 * NVLink all-gather of per-rank slices, so that each rank uploads only 1/N of it over PCIe) */
int zkb_groth16_witness_dev(zkb_ctx* ctx, const zkb_pk* pk, const uint8_t* witness, void* d_abc_out[3]);
/* as zkb_groth16_witness_dev, but every rank computes only the rows of its "layout d" block and
 * writes them straight into its own buffers (n / world elements each): the R1CS products are
 * sharded too */
int zkb_groth16_witness_layout_dev(zkb_ctx* ctx, const zkb_pk* pk, const uint8_t* witness, uint32_t low_bits,
                                   uint32_t world, uint32_t rank, void* const d_abc_local[3]);
int zkb_groth16_msm_dev(zkb_ctx* ctx, const zkb_pk* pk, const void* d_h,
                        uint8_t partial_out[ZKB_INTERMEDIATE_BYTES]);
int zkb_groth16_assemble(zkb_ctx* ctx, const zkb_pk* pk, const uint8_t sums[ZKB_INTERMEDIATE_BYTES],
                         const uint8_t r[32], const uint8_t s[32], uint8_t proof_out[ZKB_PROOF_BYTES]);
/* One proof over all ranks of ctx's communicator (collective).  The key holds this rank's point
 * range of every query, h_query in the layout the last four-step transform produces; every rank
 * passes the same full witness (host pointer: each uploads 1/world of it over PCIe and the slices
 * are all-gathered over NVLink; or a device pointer) and every rank receives the proof.  Witness
 * map: 7 four-step NTTs; MSMs: sharded by point range; the 384-byte partial sums are gathered,
 * added and assembled on the device.  timings as for zkb_groth16_prove (ms_h2d = witness staging). */
int zkb_pk_synthetic_dist(zkb_ctx* ctx, uint32_t log_n, uint64_t seed, zkb_pk** out,
                          uint8_t vk_out[ZKB_VK_BYTES]);
int zkb_groth16_prove_dist(zkb_ctx* ctx, const zkb_pk* pk, const uint8_t* witness, const uint8_t r[32],
                           const uint8_t s[32], uint8_t proof_out[ZKB_PROOF_BYTES],
                           zkb_prove_timings* timings_opt);
/* The O(1) last step on the host, no context needed: five MSM sums + the key's fixed points + (r, s)
 * -> proof.  fixed = alpha_g1 64 | beta_g1 64 | delta_g1 64 | beta_g2 128 | delta_g2 128 (canonical).
 * This is what zkb_groth16_prove / _prove_dist / _assemble run after the MSM phase (the six 254-step
 * ladders are strictly serial: a CPU core finishes one in 0.15 ms, a lone GPU lane needs 1-3 ms;
 * ctx tunable "assemble_on_device" = 1 selects the one-CTA device kernel instead, for comparison). */
int zkb_groth16_assemble_host(const uint8_t fixed[448], const uint8_t sums[ZKB_INTERMEDIATE_BYTES],
                              const uint8_t r[32], const uint8_t s[32], uint8_t proof_out[ZKB_PROOF_BYTES]);
/* CPU pairing check e(A,B) = e(alpha,beta) e(ic0 + x*ic1, gamma) e(C,delta); 0 = valid,
 * ZKB_EVERIFY = well-formed but invalid, ZKB_EINVAL = malformed (off-curve / non-canonical) */
int zkb_groth16_verify(const uint8_t vk[ZKB_VK_BYTES], const uint8_t proof[ZKB_PROOF_BYTES],
                       const uint8_t public_input[32]);
/* the same check for a key with any number of public inputs: vk_fixed = alpha_g1 64 | beta_g2 128 |
 * gamma_g2 128 | delta_g2 128, ic = n_public + 1 G1 points, public_inputs = n_public canonical Fr */
int zkb_groth16_verify_ex(const uint8_t vk_fixed[448], const uint8_t* ic, uint32_t n_public,
                          const uint8_t proof[ZKB_PROOF_BYTES], const uint8_t* public_inputs);
/* Strict well-formedness check of a proof / public input pair BEFORE it is handed to the settlement
 * layer (CPU, needs no key): the reference's parse_proof (src/settlement/ethereum/mod.rs:445-473)
 * accepts any U256, the contract then reverts (InvalidProof / NewStateRootNotInsidePrime,
 * src/settlement/ethereum/interfaces/zkvm.rs:82-130) after the L1 transaction has been paid for.
 * Returns ZKB_OK, or ZKB_EINVAL with *reasons_opt = OR of the ZKB_BAD_* bits below. */
#define ZKB_BAD_A_RANGE 0x001u     /* a coordinate of pi_a is >= q */
#define ZKB_BAD_A_CURVE 0x002u     /* pi_a is not on y^2 = x^3 + 3 */
#define ZKB_BAD_B_RANGE 0x004u
#define ZKB_BAD_B_CURVE 0x008u     /* pi_b is not on the twist y^2 = x^3 + 3/(9+u) */
#define ZKB_BAD_B_SUBGROUP 0x010u  /* on the twist but outside the order-r subgroup */
#define ZKB_BAD_C_RANGE 0x020u
#define ZKB_BAD_C_CURVE 0x040u
#define ZKB_BAD_INPUT_RANGE 0x080u /* public input >= r */
#define ZKB_BAD_INFINITY 0x100u    /* a proof element is the point at infinity (all-zero encoding) */
int zkb_proof_check(const uint8_t proof[ZKB_PROOF_BYTES], const uint8_t public_input[32],
                    uint32_t* reasons_opt);

/* test hook: the decomposed final exponentiation of the verifier against plain square-and-multiply on
 * Miller values built from the proof's points; ZKB_OK when they agree */
int zkb_selftest_final_exp(const uint8_t proof[ZKB_PROOF_BYTES]);
/* test hook: proof assembly through the per-key fixed-base tables == through plain ladders */
int zkb_selftest_assemble_tables(const uint8_t fixed[448], const uint8_t sums[ZKB_INTERMEDIATE_BYTES],
                                 const uint8_t r[32], const uint8_t s[32]);

/* ---- circuit files -------------------------------------------------------------
 * Readers for the iden3 binary formats circom / snarkjs write (.r1cs constraint systems, .wtns
 * witnesses) over BN254's scalar field; host code, no allocation.  Wire order is
 * (1, public outputs, public inputs, private inputs) = the z order of zkb_pk_desc, with
 * n_public = n_pub_out + n_pub_in.  Two-pass: inspect sizes the caller's buffers, read fills them
 * (row_ptr[k]: n_constraints + 1 entries, col[k]: nnz[k], coeff[k]: 32 * nnz[k] bytes, canonical).
 * ZKB_EINVAL for a bad magic / version / truncated section / another prime / a wire index out of
 * range / a non-canonical value / a witness whose wire 0 is not 1. */
typedef struct zkb_r1cs_info {
  uint32_t n_wires, n_pub_out, n_pub_in, n_prv_in, n_constraints;
  uint64_t n_labels;
  uint64_t nnz[3];   /* non-zeros of A, B, C */
} zkb_r1cs_info;
int zkb_r1cs_inspect(const uint8_t* file, uint64_t len, zkb_r1cs_info* out);
int zkb_r1cs_read(const uint8_t* file, uint64_t len, uint32_t* const row_ptr[3], uint32_t* const col[3],
                  uint8_t* const coeff[3]);
int zkb_wtns_inspect(const uint8_t* file, uint64_t len, uint32_t* n_witness_out);
int zkb_wtns_read(const uint8_t* file, uint64_t len, uint8_t* witness_out);
/* snarkjs / rapidsnark Groth16 proving keys (.zkey, BN254).  The container stores points as affine
 * coordinates in Montgomery form and matrix coefficients times 2^512 mod r; the reader hands back
 * canonical little-endian values (all-zero point = infinity), ready for zkb_pk_desc with
 * h_basis = 1.  Only A and B are stored (a prover takes c = a * b on the domain), A includes the
 * n_public + 1 input rows, and entries may come in any order: the reader builds CSR with n_rows =
 * highest row index + 1.  Buffers (caller-allocated from zkb_zkey_info): row_ptr[k] n_rows + 1,
 * col[k] nnz[k], coeff[k] 32 * nnz[k]; alpha_g1 / beta_g1 / delta_g1 64; beta_g2 / gamma_g2 /
 * delta_g2 128; ic 64 * (n_public + 1); a_query, b_g1_query 64 * n_vars; b_g2_query 128 * n_vars;
 * c_query 64 * (n_vars - n_public - 1); h_query 64 * 2^log_n.  Null output pointers are skipped.
 * ZKB_EINVAL for a bad magic / version / protocol / prime / truncated or mis-sized section / an
 * index out of range / a value that is not reduced.  Points are NOT checked for curve membership
 * here; zkb_pk_load does that on the device.  The layout is restated, with a writer, in the test
 * tree; it is not pinned against a file written by snarkjs itself (none exists in this build
 * environment). */
typedef struct zkb_zkey_info {
  uint32_t n_vars, n_public, log_n, n_rows;
  uint64_t nnz[2];   /* non-zeros of A, B */
} zkb_zkey_info;
typedef struct zkb_zkey_data {
  uint32_t* row_ptr[2];
  uint32_t* col[2];
  uint8_t* coeff[2];
  uint8_t *alpha_g1, *beta_g1, *delta_g1, *beta_g2, *gamma_g2, *delta_g2, *ic;
  uint8_t *a_query, *b_g1_query, *b_g2_query, *c_query, *h_query;
} zkb_zkey_data;
int zkb_zkey_inspect(const uint8_t* file, uint64_t len, zkb_zkey_info* out);
int zkb_zkey_read(const uint8_t* file, uint64_t len, const zkb_zkey_data* out);

/* ---- test / measurement hooks ------------------------------------------------------------------ */
/* element-wise field ops on host arrays of canonical elements, executed by the device field
 * library: field 0 = Fq, 1 = Fr; op 0 add, 1 sub, 2 mul, 3 sqr, 4 inverse, 5 neg,
 * 6 sqr through the dedicated (experimental, not used by the hot kernels) squaring routine,
 * 7 / 8 the fused dual product a*b + (a+b)*(a-b) / a*b - (a+b)*(a-b) (one shared reduction),
 * 9 inverse by Fermat exponentiation (op 4 is the binary extended Euclid the kernels use) */
int zkb_field_batch_op(zkb_ctx* ctx, int field, int op, const uint8_t* a, const uint8_t* b,
                       uint8_t* out, uint64_t n);
/* curve ops through the device XYZZ library: group 0 = G1, 1 = G2;
 * op 0: out[i] = a[i] + b[i]   (b = affine points)
 * op 1: out[i] = 2 a[i]
 * op 2: out[i] = k[i] * a[i]   (b = 32-byte scalars) */
int zkb_curve_batch_op(zkb_ctx* ctx, int group, int op, const uint8_t* a, const uint8_t* b,
                       uint8_t* out, uint64_t n);
/* integer-pipe microbenchmark: results[0] = IMAD.WIDE.U32 carry-chain ops/s,
 * results[1] = Montgomery multiplications/s (register resident), results[2] = SM clock (MHz)
 * measured over the run with clock64 */
int zkb_bench_int_pipe(zkb_ctx* ctx, int iters, double results[3]);
/* kernels of this library launched through ctx so far (library kernels such as the CUB sort
 * are not counted) */
uint64_t zkb_ctx_launches(const zkb_ctx* ctx);
/* per-stage device times (ms) of the last MSM run with "msm_profile" = 1:
 * digits | sort | accumulate L1 | accumulate L2 + finish | bucket reduce + window sum + Horner */
int zkb_msm_last_timings(zkb_ctx* ctx, float ms_out[5]);

#pragma GCC visibility pop
#ifdef __cplusplus
}
#endif
#endif /* ZKB_H */
