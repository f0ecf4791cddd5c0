// capi.cu — extern "C" entry points of include/zkb.h: context, NTT, MSM, test hooks.
// (The Groth16 entry points live in groth16.cu, the pairing verifier in verify.cpp.)
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <new>

#include "context.h"
#include "ec.cuh"

using namespace zkb;
using zkb::host::HFr;

namespace zkb {

int set_error(zkb_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->err = msg;
  return code;
}
int cuda_fail(zkb_ctx* ctx, cudaError_t e, const char* what) {
  std::string m = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
  cudaGetLastError();
  return set_error(ctx, e == cudaErrorMemoryAllocation ? ZKB_ENOMEM : ZKB_ECUDA, m);
}

static const uint64_t kRoot2_28[4] = {0x9bd61b6e725b19f0ull, 0x402d111e41112ed4ull,
                                      0x00e0a7eb8ef62abcull, 0x2a3c09f0a58a7e85ull};

static void hfr_to_words(const HFr& a, uint32_t out[8]) { std::memcpy(out, a.l, 32); }

// w_{2^log_n} (or its inverse) in Montgomery form
void host_fr_root(uint32_t log_n, int inverse, uint32_t out_mont[8]) {
  HFr w{{kRoot2_28[0], kRoot2_28[1], kRoot2_28[2], kRoot2_28[3]}};
  w = w.to_mont();
  for (uint32_t i = log_n; i < 28; i++) w = w.sqr();
  if (inverse) w = w.inverse();
  hfr_to_words(w, out_mont);
}

static HFr hfr_pow_u64(HFr b, uint64_t e) {
  HFr r = HFr::one();
  while (e) {
    if (e & 1) r = r * b;
    b = b.sqr();
    e >>= 1;
  }
  return r;
}

static std::vector<uint32_t> ntt_plan(uint32_t log_n) {
  std::vector<uint32_t> ks;
  if (log_n <= 8) {
    ks.push_back(log_n);
    return ks;
  }
  uint32_t P = (log_n + 7) / 8;
  uint32_t base = log_n / P, rem = log_n % P;
  for (uint32_t i = 0; i < P; i++) ks.push_back(i < rem ? base + 1 : base);
  return ks;
}

static int ensure_stage_tables(zkb_ctx* ctx) {
  if (ctx->stage_fwd.p) return ZKB_OK;
  ZKB_CUDA(ctx, ctx->stage_fwd.ensure(256 * 32));
  ZKB_CUDA(ctx, ctx->stage_inv.ensure(256 * 32));
  uint32_t one[8];
  hfr_to_words(HFr::one(), one);
  for (int inv = 0; inv < 2; inv++) {
    char* dst = reinterpret_cast<char*>(inv ? ctx->stage_inv.p : ctx->stage_fwd.p);
    for (uint32_t b = 0; b < 8; b++) {
      uint32_t root[8];
      host_fr_root(b + 1, inv, root);
      fr_powers_launch(dst + ((1u << b) - 1) * 32, root, one, 1ull << b, 1, ctx->stream);
    }
  }
  ZKB_CUDA(ctx, cudaGetLastError());
  return ZKB_OK;
}

static int get_domain(zkb_ctx* ctx, uint32_t log_n, int inverse, NttDomain** out) {
  uint64_t key = uint64_t(log_n) * 2 + (inverse ? 1 : 0);
  auto it = ctx->domains.find(key);
  if (it != ctx->domains.end()) {
    *out = it->second;
    return ZKB_OK;
  }
  NttDomain* d = new (std::nothrow) NttDomain();
  if (!d) return set_error(ctx, ZKB_ENOMEM, "out of host memory");
  d->log_n = log_n;
  d->ks = ntt_plan(log_n);
  if (d->ks.size() > 1) {
    d->lo_bits = (log_n + 1) / 2;
    uint32_t hi_bits = log_n - d->lo_bits;
    cudaError_t e;
    if ((e = d->tw_lo.ensure((1ull << d->lo_bits) * 32)) != cudaSuccess ||
        (e = d->tw_hi.ensure((1ull << hi_bits) * 32)) != cudaSuccess) {
      delete d;
      return cuda_fail(ctx, e, "twiddle allocation");
    }
    uint32_t root[8], one[8];
    host_fr_root(log_n, inverse, root);
    HFr c = HFr::one();
    if (inverse) c = hfr_pow_u64(HFr::from_u64(2), log_n).inverse();   // 1/n folded into pass 0
    hfr_to_words(HFr::one(), one);
    uint32_t cw[8];
    hfr_to_words(c, cw);
    fr_powers_launch(d->tw_lo.p, root, one, 1ull << d->lo_bits, 1, ctx->stream);
    fr_powers_launch(d->tw_hi.p, root, one, 1ull << hi_bits, 1ull << d->lo_bits, ctx->stream);
    if (inverse) {
      if ((e = d->tw_hi_ninv.ensure((1ull << hi_bits) * 32)) != cudaSuccess) {
        delete d;
        return cuda_fail(ctx, e, "twiddle allocation");
      }
      fr_powers_launch(d->tw_hi_ninv.p, root, cw, 1ull << hi_bits, 1ull << d->lo_bits,
                       ctx->stream);
    }
  }
  ctx->domains[key] = d;
  *out = d;
  return ZKB_OK;
}

static int ensure_full_tables(zkb_ctx* ctx, NttDomain* d, int inverse) {
  if (d->full_ready) return ZKB_OK;
  size_t P = d->ks.size();
  d->full.resize(P > 0 ? P - 1 : 0);
  uint64_t M = 1ull << d->log_n;
  for (size_t p = 0; p + 1 < P; p++) {
    uint32_t k = d->ks[p];
    uint32_t log_m = 0;
    while ((1ull << log_m) < M) log_m++;
    ZKB_CUDA(ctx, d->full[p].ensure(M * 32));
    const void* hi = (p == 0 && inverse) ? d->tw_hi_ninv.p : d->tw_hi.p;
    fr_twiddle_full_launch(d->full[p].p, d->tw_lo.p, hi, d->lo_bits, d->log_n - log_m, log_m - k,
                           M, ctx->stream);
    M >>= k;
  }
  ZKB_CUDA(ctx, cudaGetLastError());
  d->full_ready = true;
  return ZKB_OK;
}

int coset_tables(zkb_ctx* ctx, uint32_t log_n, CosetTables** out) {
  auto it = ctx->cosets.find(log_n);
  if (it != ctx->cosets.end()) {
    *out = it->second;
    return ZKB_OK;
  }
  CosetTables* t = new (std::nothrow) CosetTables();
  if (!t) return set_error(ctx, ZKB_ENOMEM, "out of host memory");
  uint64_t n = 1ull << log_n;
  bool single = ntt_plan(log_n).size() == 1;   // then 1/n lives in the post tables
  HFr g = HFr::from_u64(5);
  HFr gi = g.inverse();
  HFr ninv = hfr_pow_u64(HFr::from_u64(2), log_n).inverse();
  HFr zinv = (hfr_pow_u64(g, n) - HFr::one()).inverse();
  HFr one = HFr::one();
  uint32_t gw[8], giw[8], onew[8], c_inv[8], c_invz[8], ninvw[8];
  hfr_to_words(g, gw);
  hfr_to_words(gi, giw);
  hfr_to_words(one, onew);
  hfr_to_words(single ? ninv : one, c_inv);
  hfr_to_words(single ? ninv * zinv : zinv, c_invz);
  hfr_to_words(ninv, ninvw);
  cudaError_t e;
  if ((e = t->fwd.ensure(n * 32)) != cudaSuccess || (e = t->inv.ensure(n * 32)) != cudaSuccess ||
      (e = t->inv_z.ensure(n * 32)) != cudaSuccess ||
      (single && (e = t->ninv.ensure(n * 32)) != cudaSuccess)) {
    t->fwd.release(); t->inv.release(); t->inv_z.release(); t->ninv.release();
    delete t;
    return cuda_fail(ctx, e, "coset table allocation");
  }
  fr_powers_launch(t->fwd.p, gw, onew, n, 1, ctx->stream);
  fr_powers_launch(t->inv.p, giw, c_inv, n, 1, ctx->stream);
  fr_powers_launch(t->inv_z.p, giw, c_invz, n, 1, ctx->stream);
  if (single) fr_powers_launch(t->ninv.p, onew, ninvw, n, 1, ctx->stream);
  ZKB_CUDA(ctx, cudaGetLastError());
  ctx->cosets[log_n] = t;
  *out = t;
  return ZKB_OK;
}

int ntt_run(zkb_ctx* ctx, const void* d_in, void* d_out, uint32_t log_n, int inverse,
            const void* pre_scale, const void* post_scale, int skip_ninv, int mont_in,
            int mont_out, uint64_t batch, const NttRemap* remap) {
  if (batch == 0) return ZKB_OK;
  if (remap && remap->in_ilv && (batch & (batch - 1)))
    return set_error(ctx, ZKB_EINVAL, "element-major input needs a power-of-two batch");
  if (log_n < 1 || log_n > 28) return set_error(ctx, ZKB_EINVAL, "log_n must be in [1, 28]");
  int rc = ensure_stage_tables(ctx);
  if (rc) return rc;
  NttDomain* d;
  if ((rc = get_domain(ctx, log_n, inverse, &d))) return rc;
  const size_t P = d->ks.size();
  const uint64_t n = 1ull << log_n;
  const uint64_t total = n * batch;
  if ((total >> 11) * 8 >= (1ull << 31)) return set_error(ctx, ZKB_EINVAL, "batch too large");
  const bool use_full = P > 1 && ctx->ntt_full_twiddles && !(inverse && skip_ninv);   // full tables fold 1/n
  if (use_full && (rc = ensure_full_tables(ctx, d, inverse))) return rc;
  if (inverse && P == 1 && !post_scale && !skip_ninv) {
    if (batch != 1) return set_error(ctx, ZKB_EINVAL, "batched single-pass inverse needs an explicit scale table");
    CosetTables* t;
    if ((rc = coset_tables(ctx, log_n, &t))) return rc;
    post_scale = t->ninv.p;
  }
  if (!ctx->msm.flag.p) {
    ZKB_CUDA(ctx, ctx->msm.flag.ensure(16));
    ZKB_CUDA(ctx, cudaMemsetAsync(ctx->msm.flag.p, 0, 16, ctx->stream));
  }
  NttPassArgs a{};
  a.xform_log = log_n;
  if (remap && remap->in_ilv) {
    a.in_ilv_log = 0;
    while ((1ull << a.in_ilv_log) < batch) a.in_ilv_log++;
  }
  a.flag = (uint32_t*)ctx->msm.flag.p;
  a.stage_tw = inverse ? ctx->stage_inv.p : ctx->stage_fwd.p;
  a.tw_lo = d->tw_lo.p;
  a.tw_lo_bits = d->lo_bits;
  a.tw_lo_mask = (1ull << d->lo_bits) - 1;
  void* work = d_out;
  if (P > 1) {
    ZKB_CUDA(ctx, ctx->ntt_tmp.ensure(total * 32));
    work = ctx->ntt_tmp.p;
  }
  uint64_t M = n;
  uint32_t log_m = log_n;
  for (size_t p = 0; p + 1 < P; p++) {
    uint32_t k = d->ks[p];
    a.in = p == 0 ? d_in : work;
    a.in_ilv = (p == 0 && remap) ? remap->in_ilv : 0;
    a.out = work;
    a.k = k;
    a.M = M;
    a.S = uint32_t(M >> k);
    a.tw_shift = log_n - log_m;
    a.tw_hi = (p == 0 && inverse && !skip_ninv) ? d->tw_hi_ninv.p : d->tw_hi.p;
    a.tw_full = (use_full && d->full_ready) ? d->full[p].p : nullptr;
    a.pre_scale = p == 0 ? pre_scale : nullptr;
    a.post_scale = nullptr;
    a.mont_in = p == 0 ? mont_in : 1;
    a.mont_out = 1;
    uint32_t CT = 8u << (8 - k);
    uint32_t tiles = uint32_t(total / M) * ((a.S + CT - 1) / CT);
    ntt_launch_pass(a, false, tiles, ctx->stream);
    M >>= k;
    log_m -= k;
  }
  {
    uint32_t k = d->ks[P - 1];
    a.in = P == 1 ? d_in : work;
    a.in_ilv = (P == 1 && remap) ? remap->in_ilv : 0;
    if (remap && remap->out_win) {
      a.out_win = 1;
      a.out_col_bits = remap->out_col_bits;
      for (int i = 0; i < 16; i++) a.out_wins[i] = remap->out_wins[i];
    }
    a.out = d_out;
    a.k = k;
    a.tw_full = nullptr;
    a.pre_scale = P == 1 ? pre_scale : nullptr;
    a.post_scale = post_scale;
    a.mont_in = P == 1 ? mont_in : 1;
    a.mont_out = mont_out;
    a.row_bits = log_n - k;
    a.first_digit_bits = P > 1 ? d->ks[0] : 0;
    a.n_row_digits = uint32_t(P - 1);
    for (size_t i = 0; i + 1 < P; i++) a.row_digit_bits[i] = d->ks[i];
    uint32_t CT = 8u << (8 - k);
    uint32_t k0_range = 1u << a.first_digit_bits;
    uint32_t tiles = ((k0_range + CT - 1) / CT) << (a.row_bits - a.first_digit_bits);
    a.batch_log = log_n;
    a.tiles_per_batch = batch > 1 ? tiles : 0;
    uint64_t n_tiles = uint64_t(tiles) * batch;
    if (P == 1 && batch > 1) {          // columns of a tile = consecutive transforms of the batch
      a.batch_cols = 1;
      a.batch_count = uint32_t(batch);
      n_tiles = (batch + CT - 1) / CT;
    }
    ntt_launch_pass(a, true, uint32_t(n_tiles), ctx->stream);
  }
  ZKB_CUDA(ctx, cudaGetLastError());
  return ZKB_OK;
}

// ---------------------------------------------------------------- MSM glue
static int ensure_msm_buffers(zkb_ctx* ctx, const MsmPlan& p, MsmBuffers* b, bool second_set = false) {
  MsmScratch& s = ctx->msm;
  if (!s.flag.p) {
    ZKB_CUDA(ctx, s.flag.ensure(16));
    ZKB_CUDA(ctx, cudaMemsetAsync(s.flag.p, 0, 16, ctx->stream));
  }
  if (second_set) ZKB_CUDA(ctx, s.buckets2.ensure(p.n_buckets * p.point_bytes));
  b->buckets2 = s.buckets2.p;
  b->flag = (uint32_t*)s.flag.p;
  ZKB_CUDA(ctx, s.keys_in.ensure(p.m * 8));    // 8 B per entry: the bucket sort's intermediate (key, value) array
  ZKB_CUDA(ctx, s.vals_in.ensure(p.m * 4));
  ZKB_CUDA(ctx, s.keys_out.ensure(p.m * 4));
  ZKB_CUDA(ctx, s.vals_out.ensure(p.m * 4));
  ZKB_CUDA(ctx, s.sort_temp.ensure(p.sort_temp_bytes + 16));
  ZKB_CUDA(ctx, s.m_valid.ensure(8));
  ZKB_CUDA(ctx, s.buckets.ensure(p.n_buckets * p.point_bytes));
  ZKB_CUDA(ctx, s.part1_keys.ensure(2ull * p.l1_threads * 4));
  ZKB_CUDA(ctx, s.part1_pts.ensure(2ull * p.l1_threads * p.point_bytes));
  ZKB_CUDA(ctx, s.part2_keys.ensure(2ull * p.l2_threads * 4));
  ZKB_CUDA(ctx, s.part2_pts.ensure(2ull * p.l2_threads * p.point_bytes));
  ZKB_CUDA(ctx, s.group_sums.ensure(msm_tail_scratch_points(p) * p.point_bytes));
  ZKB_CUDA(ctx, s.window_sums.ensure(uint64_t(p.red_levels + 2) * p.n_sets * p.point_bytes));
  ZKB_CUDA(ctx, s.result_xyzz.ensure(p.point_bytes));
  ZKB_CUDA(ctx, s.result_affine.ensure(p.point_bytes / 2));
  b->keys_in = (uint32_t*)s.keys_in.p;
  b->vals_in = (uint32_t*)s.vals_in.p;
  b->keys_out = (uint32_t*)s.keys_out.p;
  b->vals_out = (uint32_t*)s.vals_out.p;
  b->sort_temp = s.sort_temp.p;
  b->m_valid = (uint64_t*)s.m_valid.p;
  b->buckets = s.buckets.p;
  b->part1_keys = (uint32_t*)s.part1_keys.p;
  b->part1_pts = s.part1_pts.p;
  b->part2_keys = (uint32_t*)s.part2_keys.p;
  b->part2_pts = s.part2_pts.p;
  b->group_sums = s.group_sums.p;
  b->window_sums = s.window_sums.p;
  b->result_xyzz = s.result_xyzz.p;
  b->result_affine = s.result_affine.p;
  b->ev = nullptr;
  return ZKB_OK;
}

static int msm_prepare(zkb_ctx* ctx, const zkb_bases* b, uint64_t first, uint64_t n) {
  if (!b || b->device != ctx->device) return set_error(ctx, ZKB_EINVAL, "bases belong to another device");
  if (first > b->n || n > b->n - first) return set_error(ctx, ZKB_EINVAL, "MSM range exceeds the bases");
  if (b->n * (b->precomp_c ? b->precomp_nwin : 1) >= (1ull << 31))
    return set_error(ctx, ZKB_EINVAL, "MSM larger than 2^31 - 1 (expanded) points");
  return ZKB_OK;
}

static int msm_profile_events(zkb_ctx* ctx, MsmBuffers* buf) {
  ctx->msm_ev_valid = false;
  if (ctx->msm_profile) {
    for (int i = 0; i < 6; i++)
      if (!ctx->msm_ev[i]) ZKB_CUDA(ctx, cudaEventCreate(&ctx->msm_ev[i]));
    buf->ev = ctx->msm_ev;
    ctx->msm_ev_valid = true;
  }
  return ZKB_OK;
}

int msm_exec(zkb_ctx* ctx, const zkb_bases* b, uint64_t first, const void* d_scalars, uint64_t n,
             void* d_out_affine) {
  int rc = msm_prepare(ctx, b, first, n);
  if (rc) return rc;
  const size_t aff = b->is_g2 ? 128 : 64;
  if (n == 0) {
    ZKB_CUDA(ctx, cudaMemsetAsync(d_out_affine, 0, aff, ctx->stream));
    return ZKB_OK;
  }
  MsmPlan p = b->precomp_c ? msm_make_plan(n, b->is_g2, b->precomp_c, ctx->sm_count, b->n)
                           : msm_make_plan(n, b->is_g2, int(ctx->msm_window), ctx->sm_count);
  MsmBuffers buf;
  if ((rc = ensure_msm_buffers(ctx, p, &buf))) return rc;
  buf.result_affine = d_out_affine;
  if ((rc = msm_profile_events(ctx, &buf))) return rc;
  // values index the whole base array: the sub-range starts at `first`
  msm_digits_and_sort(p, buf, d_scalars, n, uint32_t(first), ctx->stream);
  if (b->is_g2) {
    msm_accumulate_g2(p, buf, b->d_pts, buf.buckets, ctx->stream);
    msm_tail_g2(p, buf, ctx->stream);
  } else {
    msm_accumulate_g1(p, buf, b->d_pts, buf.buckets, ctx->stream);
    msm_tail_g1(p, buf, ctx->stream);
  }
  ZKB_CUDA(ctx, cudaGetLastError());
  return ZKB_OK;
}

// k MSMs over the SAME scalars (the A / B1 / B2 / L MSMs of a Groth16 proof all use the witness):
// digits and sort run once, accumulate + tail once per base set.  All base sets must have the same
// length and window layout.  ev_opt[j] (optional) is recorded after MSM j.
int msm_exec_multi(zkb_ctx* ctx, const zkb_bases* const* bs, int k, const void* d_scalars, uint64_t first,
                   uint64_t n, void* const* d_outs, cudaEvent_t* ev_opt) {
  if (k <= 0) return ZKB_OK;
  for (int j = 0; j < k; j++) {
    int rc = msm_prepare(ctx, bs[j], first, n);
    if (rc) return rc;
    if (bs[j]->n != bs[0]->n || bs[j]->precomp_c != bs[0]->precomp_c)
      return set_error(ctx, ZKB_EINVAL, "shared-scalar MSMs need base sets of one shape");
  }
  if (n == 0) {
    for (int j = 0; j < k; j++) {
      ZKB_CUDA(ctx, cudaMemsetAsync(d_outs[j], 0, bs[j]->is_g2 ? 128 : 64, ctx->stream));
      if (ev_opt) cudaEventRecord(ev_opt[j], ctx->stream);
    }
    return ZKB_OK;
  }
  const zkb_bases* b0 = bs[0];
  int any_g2 = 0;
  for (int j = 0; j < k; j++) any_g2 |= bs[j]->is_g2;
  // buffers sized for the widest point type; the chunk walk uses each group's own thread count
  MsmPlan p1 = b0->precomp_c ? msm_make_plan(n, 0, b0->precomp_c, ctx->sm_count, b0->n)
                             : msm_make_plan(n, 0, int(ctx->msm_window), ctx->sm_count);
  MsmPlan p2 = msm_make_plan(n, 1, p1.c, ctx->sm_count, p1.val_stride);
  MsmBuffers buf;
  int rc;
  if (any_g2 && (rc = ensure_msm_buffers(ctx, p2, &buf))) return rc;
  if ((rc = ensure_msm_buffers(ctx, p1, &buf))) return rc;
  msm_digits_and_sort(p1, buf, d_scalars, n, uint32_t(first), ctx->stream);
  for (int j = 0; j < k; j++) {
    buf.result_affine = d_outs[j];
    if (bs[j]->is_g2) {
      msm_accumulate_g2(p2, buf, bs[j]->d_pts, buf.buckets, ctx->stream);
      msm_tail_g2(p2, buf, ctx->stream);
    } else {
      msm_accumulate_g1(p1, buf, bs[j]->d_pts, buf.buckets, ctx->stream);
      msm_tail_g1(p1, buf, ctx->stream);
    }
    if (ev_opt) cudaEventRecord(ev_opt[j], ctx->stream);
  }
  ZKB_CUDA(ctx, cudaGetLastError());
  return ZKB_OK;
}

// The MSM phase of one Groth16 proof.  qs = a, b_g1, l (G1) and b_g2 (G2) over the witness z (one
// digit extraction + one sort for the four of them), h_query over h.  When every G1 query carries
// window multiples of one width, the four G1 walks fill four bucket sets that ONE tail reduces (the
// tail is a chain of short latency-bound launches: run four times it costs four times as much, run
// once over four sets it costs the same as for one).  res: msm_a 64 | msm_b1 64 | msm_b2 128 | msm_l 64 | msm_h 64.
int msm_exec_prove(zkb_ctx* ctx, const zkb_bases* a, const zkb_bases* b1, const zkb_bases* l, const zkb_bases* b2,
                   const zkb_bases* h, const void* d_z, uint64_t nz, const void* d_h, uint64_t nh, void* d_res,
                   cudaEvent_t* ev4_opt) {
  // ev4_opt (optional): recorded after the a, b_g1, l and b_g2 walks
  char* res = (char*)d_res;
  const bool batched = a->precomp_c && a->precomp_c == b1->precomp_c && a->precomp_c == l->precomp_c &&
                       a->precomp_c == b2->precomp_c && a->precomp_c == h->precomp_c && a->n == b1->n &&
                       a->n == l->n && a->n == b2->n && nz && nh;
  if (!batched) {
    const zkb_bases* qs[4] = {a, b1, l, b2};
    void* outs4[4] = {res, res + 64, res + 256, res + 128};
    cudaEvent_t evs[4];
    if (ev4_opt) { evs[0] = ev4_opt[0]; evs[1] = ev4_opt[1]; evs[2] = ev4_opt[2]; evs[3] = ev4_opt[3]; }
    int rc = msm_exec_multi(ctx, qs, 4, d_z, 0, nz, outs4, ev4_opt ? evs : nullptr);
    if (rc) return rc;
    return msm_exec(ctx, h, 0, d_h, nh, res + 320);
  }
  int rc;
  const zkb_bases* all[5] = {a, b1, l, b2, h};
  for (int j = 0; j < 5; j++)
    if ((rc = msm_prepare(ctx, all[j], 0, j == 4 ? nh : nz))) return rc;
  const int c = a->precomp_c;
  MsmPlan p1 = msm_make_plan(nz, 0, c, ctx->sm_count, a->n);
  MsmPlan p2 = msm_make_plan(nz, 1, c, ctx->sm_count, b2->n);
  MsmPlan ph = msm_make_plan(nh, 0, c, ctx->sm_count, h->n);
  MsmPlan pt = p1;                         // the shared tail: four independent bucket sets
  pt.n_sets = 4;
  pt.n_buckets = 4ull * p1.half;
  pt.red_run_total = 4 * p1.red_run_total;
  MsmBuffers buf;
  if ((rc = ensure_msm_buffers(ctx, p2, &buf))) return rc;
  if ((rc = ensure_msm_buffers(ctx, ph, &buf))) return rc;
  if ((rc = ensure_msm_buffers(ctx, p1, &buf))) return rc;
  if ((rc = ensure_msm_buffers(ctx, pt, &buf, false))) return rc;
  MsmScratch& sc = ctx->msm;
  ZKB_CUDA(ctx, sc.buckets2.ensure(p2.n_buckets * p2.point_bytes));
  // Order on the main stream: sort(z), the four witness walks, sort(h), the H walk — each walk is
  // one resident wave that owns every register of the GPU, nothing overlaps it.  What follows is
  // latency-bound (partial merges, bucket reductions, window sums, Horner: small grids of dependent
  // point additions) and runs side by side: one stream per MSM for its partial merges, then the
  // shared G1 tail on the main stream next to the G2 tail on its own.
  const MsmPlan* plans[5] = {&p1, &p1, &p1, &p2, &ph};
  size_t part_off[6] = {0};
  for (int j = 0; j < 5; j++) {
    const MsmPlan& q = *plans[j];
    const size_t bytes = 2ull * q.l1_threads * (4 + q.point_bytes) + 2ull * q.l2_threads * (4 + q.point_bytes) + 1024;
    part_off[j + 1] = part_off[j] + ((bytes + 255) & ~size_t(255));
  }
  ZKB_CUDA(ctx, sc.multi_part.ensure(part_off[5]));
  ZKB_CUDA(ctx, sc.group_sums2.ensure(msm_tail_scratch_points(p2) * p2.point_bytes));
  ZKB_CUDA(ctx, sc.window_sums2.ensure(uint64_t(p2.red_levels + 2) * p2.n_sets * p2.point_bytes));
  for (int j = 0; j < 5; j++)
    if (!ctx->side[j]) ZKB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->side[j], cudaStreamNonBlocking));
  for (int j = 0; j < 6; j++)
    if (!ctx->side_ev[j]) ZKB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->side_ev[j], cudaEventDisableTiming));
  MsmBuffers bj[5];
  for (int j = 0; j < 5; j++) {
    const MsmPlan& q = *plans[j];
    bj[j] = buf;
    bj[j].ev = nullptr;
    char* base = (char*)sc.multi_part.p + part_off[j];
    bj[j].part1_keys = (uint32_t*)base;
    base += (2ull * q.l1_threads * 4 + 255) & ~size_t(255);
    bj[j].part1_pts = base;
    base += 2ull * q.l1_threads * q.point_bytes;
    bj[j].part2_keys = (uint32_t*)base;
    base += (2ull * q.l2_threads * 4 + 255) & ~size_t(255);
    bj[j].part2_pts = base;
  }
  cudaStream_t st = ctx->stream;
  char* g1b = (char*)buf.buckets;
  const size_t set_bytes = size_t(p1.half) * p1.point_bytes;
  void* targets[5] = {g1b, g1b + set_bytes, g1b + 2 * set_bytes, sc.buckets2.p, g1b + 3 * set_bytes};
  msm_digits_and_sort(p1, buf, d_z, nz, 0, st);
  msm_walk_g1(p1, bj[0], a->d_pts, targets[0], st);
  if (ev4_opt) cudaEventRecord(ev4_opt[0], st);
  msm_walk_g1(p1, bj[1], b1->d_pts, targets[1], st);
  if (ev4_opt) cudaEventRecord(ev4_opt[1], st);
  msm_walk_g1(p1, bj[2], l->d_pts, targets[2], st);
  if (ev4_opt) cudaEventRecord(ev4_opt[2], st);
  msm_walk_g2(p2, bj[3], b2->d_pts, targets[3], st);
  if (ev4_opt) cudaEventRecord(ev4_opt[3], st);
  msm_digits_and_sort(ph, buf, d_h, nh, 0, st);
  msm_walk_g1(ph, bj[4], h->d_pts, targets[4], st);
  ZKB_CUDA(ctx, cudaEventRecord(ctx->side_ev[5], st));
  for (int j = 0; j < 5; j++) {
    ZKB_CUDA(ctx, cudaStreamWaitEvent(ctx->side[j], ctx->side_ev[5], 0));
    if (j == 3) {
      msm_post_g2(p2, bj[3], targets[3], ctx->side[3]);
      MsmBuffers buf2 = buf;                 // the G2 tail right behind its merges, on its own scratch
      buf2.ev = nullptr;
      buf2.buckets = sc.buckets2.p;
      buf2.group_sums = sc.group_sums2.p;
      buf2.window_sums = sc.window_sums2.p;
      buf2.result_affine = res + 128;
      msm_tail_g2(p2, buf2, ctx->side[3]);
    } else {
      msm_post_g1(*plans[j], bj[j], targets[j], ctx->side[j]);
    }
    ZKB_CUDA(ctx, cudaEventRecord(ctx->side_ev[j], ctx->side[j]));
  }
  for (int j = 0; j < 5; j++)
    if (j != 3) ZKB_CUDA(ctx, cudaStreamWaitEvent(st, ctx->side_ev[j], 0));
  void* outs[4] = {res, res + 64, res + 256, res + 320};
  MsmBuffers buft = buf;
  buft.ev = nullptr;
  msm_tail_multi_g1(pt, buft, outs, 4, st);
  ZKB_CUDA(ctx, cudaStreamWaitEvent(st, ctx->side_ev[3], 0));
  ZKB_CUDA(ctx, cudaGetLastError());
  return ZKB_OK;
}

int msm_exec_host(zkb_ctx* ctx, const zkb_bases* b, uint64_t first, const uint8_t* h_scalars,
                  uint64_t n, void* d_out_affine) {
  int rc = msm_prepare(ctx, b, first, n);
  if (rc) return rc;
  const size_t aff = b->is_g2 ? 128 : 64;
  if (n == 0) {
    ZKB_CUDA(ctx, cudaMemsetAsync(d_out_affine, 0, aff, ctx->stream));
    return ZKB_OK;
  }
  uint32_t chunks = ctx->msm_host_chunks > 0 ? uint32_t(ctx->msm_host_chunks) : (n >= (1ull << 20) ? 3 : 1);
  if (chunks > 16) chunks = 16;
  // chunk sizes grow geometrically (1 : 2 : 4 ...): only the first, small upload is exposed, every
  // later upload hides behind the compute of the chunks before it
  uint64_t bound[17];
  bound[0] = 0;
  for (uint32_t k = 1; k <= chunks; k++) {
    const uint64_t num = (1ull << k) - 1, den = (1ull << chunks) - 1;
    bound[k] = k == chunks ? n : (n / den) * num;
  }
  // one window width for the whole MSM so that every chunk fills the same bucket array
  const uint64_t stride = b->precomp_c ? b->n : 0;
  MsmPlan whole = b->precomp_c ? msm_make_plan(n, b->is_g2, b->precomp_c, ctx->sm_count, stride)
                               : msm_make_plan(n, b->is_g2, int(ctx->msm_window), ctx->sm_count);
  uint64_t per = 0;     // the largest chunk sizes the scratch buffers
  for (uint32_t k = 0; k < chunks; k++) per = bound[k + 1] - bound[k] > per ? bound[k + 1] - bound[k] : per;
  MsmPlan p = msm_make_plan(per, b->is_g2, whole.c, ctx->sm_count, stride);
  MsmBuffers buf;
  if ((rc = ensure_msm_buffers(ctx, p, &buf, chunks > 1))) return rc;
  buf.result_affine = d_out_affine;
  ZKB_CUDA(ctx, cudaMemsetAsync(buf.flag, 0, 16, ctx->stream));   // flags raised by earlier *_dev calls are stale
  ZKB_CUDA(ctx, ctx->io_a.ensure(n * 32 + 32));
  if (!ctx->copy_stream) ZKB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  for (uint32_t k = 0; k < chunks; k++)
    if (!ctx->copy_ev[k]) ZKB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->copy_ev[k], cudaEventDisableTiming));
  // the copy stream must not overwrite io_a while earlier work on the main stream still reads it
  ZKB_CUDA(ctx, cudaEventRecord(ctx->copy_ev[0], ctx->stream));
  ZKB_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_ev[0], 0));
  for (uint32_t k = 0; k < chunks; k++) {
    const uint64_t off = bound[k], cnt = bound[k + 1] - bound[k];
    if (cnt)
      ZKB_CUDA(ctx, cudaMemcpyAsync((char*)ctx->io_a.p + off * 32, h_scalars + off * 32, cnt * 32,
                                  cudaMemcpyHostToDevice, ctx->copy_stream));
    ZKB_CUDA(ctx, cudaEventRecord(ctx->copy_ev[k], ctx->copy_stream));
  }
  bool first_done = false;
  for (uint32_t k = 0; k < chunks; k++) {
    const uint64_t off = bound[k], cnt = bound[k + 1] - bound[k];
    if (!cnt) continue;
    MsmPlan pk = cnt == per ? p : msm_make_plan(cnt, b->is_g2, whole.c, ctx->sm_count, stride);
    pk.sort_temp_bytes = p.sort_temp_bytes;
    ZKB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->copy_ev[k], 0));
    msm_digits_and_sort(pk, buf, (char*)ctx->io_a.p + off * 32, cnt, uint32_t(first + off), ctx->stream);
    void* target = !first_done ? buf.buckets : buf.buckets2;
    if (b->is_g2) msm_accumulate_g2(pk, buf, b->d_pts, target, ctx->stream);
    else msm_accumulate_g1(pk, buf, b->d_pts, target, ctx->stream);
    const bool merge = first_done;
    first_done = true;
    if (merge) {
      if (b->is_g2) msm_merge_g2(pk, buf.buckets, buf.buckets2, ctx->stream);
      else msm_merge_g1(pk, buf.buckets, buf.buckets2, ctx->stream);
    }
  }
  if (b->is_g2) msm_tail_g2(p, buf, ctx->stream);
  else msm_tail_g1(p, buf, ctx->stream);
  ZKB_CUDA(ctx, cudaGetLastError());
  return ZKB_OK;
}

// ---------------------------------------------------------------- test / setup kernels
template <class F>
__global__ void field_batch_kernel(int op, const F* a, const F* b, F* out, uint64_t n) {
  uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  F x = F::load(a + i).to_mont();
  F y = b ? F::load(b + i).to_mont() : F::zero();
  F r;
  switch (op) {
    case 0: r = x + y; break;
    case 1: r = x - y; break;
    case 2: r = x * y; break;
    case 3: r = x.sqr(); break;
    case 4: r = x.inverse(); break;
    case 6: r = x.sqr_dedicated(); break;
    case 7: r = F::mul_add_mul(x, y, x + y, x - y); break;   // fused dual product, operands up to 2p
    case 8: r = F::mul_sub_mul(x, y, x + y, x - y); break;
    case 9: r = x.inverse_fermat(); break;
    default: r = x.neg(); break;
  }
  r.from_mont().store(out + i);
}

template <class F>
__global__ void curve_batch_kernel(int op, const Affine<F>* a, const void* b, Affine<F>* out,
                                   uint64_t n) {
  uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  Affine<F> p = Affine<F>::load(a + i);
  p.x = p.x.to_mont();
  p.y = p.y.to_mont();
  XYZZ<F> acc = XYZZ<F>::from_affine(p);
  if (op == 0) {
    Affine<F> q = Affine<F>::load(reinterpret_cast<const Affine<F>*>(b) + i);
    q.x = q.x.to_mont();
    q.y = q.y.to_mont();
    acc.madd(q);
  } else if (op == 1) {
    acc = acc.dbl();
  } else {
    uint32_t k[8];
    const uint32_t* kp = reinterpret_cast<const uint32_t*>(b) + 8 * i;
#pragma unroll
    for (int j = 0; j < 8; j++) k[j] = kp[j];
    acc = mul_scalar(acc, k);
  }
  Affine<F> r = acc.to_affine();
  r.x = r.x.from_mont();
  r.y = r.y.from_mont();
  r.store(out + i);
}

// out[i] = scalars[i] * G  (affine, Montgomery) — synthetic bases and proving-key queries (K16)
template <class F>
__global__ void __launch_bounds__(128) fixed_base_mul_kernel(const uint32_t* scalars, Affine<F> gen,
                                                             Affine<F>* out, uint64_t n) {
  uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  uint32_t k[8];
#pragma unroll
  for (int j = 0; j < 8; j++) k[j] = scalars[8 * i + j];
  XYZZ<F> acc = mul_scalar(XYZZ<F>::from_affine(gen), k);
  acc.to_affine().store(out + i);
}

template <class F>
__global__ void bases_from_mont_kernel(const Affine<F>* in, Affine<F>* out, uint64_t n) {
  uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  Affine<F> p = Affine<F>::load(in + i);
  p.x = p.x.from_mont();
  p.y = p.y.from_mont();
  p.store(out + i);
}

static Affine<Fq> g1_generator_mont() {
  // (1, 2) in Montgomery form: R mod q and 2R mod q
  host::HFq one = host::HFq::one(), two = one + one;
  Affine<Fq> g;
  std::memcpy(g.x.v, one.l, 32);
  std::memcpy(g.y.v, two.l, 32);
  return g;
}
static Affine<Fq2> g2_generator_mont() {
  static const uint64_t c[4][4] = {
      {0x46debd5cd992f6edull, 0x674322d4f75edaddull, 0x426a00665e5c4479ull, 0x1800deef121f1e76ull},
      {0x97e485b7aef312c2ull, 0xf1aa493335a9e712ull, 0x7260bfb731fb5d25ull, 0x198e9393920d483aull},
      {0x4ce6cc0166fa7daaull, 0xe3d1e7690c43d37bull, 0x4aab71808dcb408full, 0x12c85ea5db8c6debull},
      {0x55acdadcd122975bull, 0xbc4b313370b38ef3ull, 0xec9e99ad690c3395ull, 0x090689d0585ff075ull}};
  host::HFq v[4];
  for (int i = 0; i < 4; i++) v[i] = host::HFq{{c[i][0], c[i][1], c[i][2], c[i][3]}}.to_mont();
  Affine<Fq2> g;
  std::memcpy(g.x.c0.v, v[0].l, 32);
  std::memcpy(g.x.c1.v, v[1].l, 32);
  std::memcpy(g.y.c0.v, v[2].l, 32);
  std::memcpy(g.y.c1.v, v[3].l, 32);
  return g;
}

int bases_from_scalars(zkb_ctx* ctx, const void* d_scalars, uint64_t n, int is_g2, void* d_out) {
  uint32_t blocks = uint32_t((n + 127) / 128);
  if (!blocks) return ZKB_OK;
  if (is_g2)
    fixed_base_mul_kernel<Fq2><<<blocks, 128, 0, ctx->stream>>>(
        reinterpret_cast<const uint32_t*>(d_scalars), g2_generator_mont(),
        reinterpret_cast<Affine<Fq2>*>(d_out), n);
  else
    fixed_base_mul_kernel<Fq><<<blocks, 128, 0, ctx->stream>>>(
        reinterpret_cast<const uint32_t*>(d_scalars), g1_generator_mont(),
        reinterpret_cast<Affine<Fq>*>(d_out), n);
  count_launches(1);
  ZKB_CUDA(ctx, cudaGetLastError());
  return ZKB_OK;
}

// canonical affine bytes already in b->d_pts -> Montgomery, with the on-curve check; synchronises
int bases_to_mont_checked(zkb_ctx* ctx, zkb_bases* b) {
  ZKB_CUDA(ctx, ctx->msm.flag.ensure(16));
  ZKB_CUDA(ctx, cudaMemsetAsync(ctx->msm.flag.p, 0, 16, ctx->stream));
  uint32_t bcoef[16];
  host_curve_b(b->is_g2, bcoef);
  if (b->is_g2) msm_bases_to_mont_g2(b->d_pts, b->n, bcoef, (uint32_t*)ctx->msm.flag.p, ctx->stream);
  else msm_bases_to_mont_g1(b->d_pts, b->n, bcoef, (uint32_t*)ctx->msm.flag.p, ctx->stream);
  uint32_t flag = 0;
  ZKB_CUDA(ctx, cudaMemcpyAsync(&flag, ctx->msm.flag.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
  ZKB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ZKB_CUDA(ctx, cudaGetLastError());
  if (flag & 2u) {
    ZKB_CUDA(ctx, cudaMemsetAsync(ctx->msm.flag.p, 0, 16, ctx->stream));
    return set_error(ctx, ZKB_EINVAL, b->is_g2 ? "a G2 base is not on the twist y^2 = x^3 + 3/(9+u)"
                                               : "a G1 base is not on the curve y^2 = x^3 + 3");
  }
  return ZKB_OK;
}

int bases_precompute_nolock(zkb_ctx* ctx, zkb_bases* b, int c) {
  if (!b || b->device != ctx->device) return set_error(ctx, ZKB_EINVAL, "bases belong to another device");
  if (b->precomp_c) return set_error(ctx, ZKB_EINVAL, "bases already carry precomputed multiples");
  if (b->n == 0) return ZKB_OK;
  if (c <= 0) c = msm_auto_window(b->n, true);
  if (c < 2 || c > 24) return set_error(ctx, ZKB_EINVAL, "window width must be in [2, 24]");
  uint32_t nwin = (255 + c - 1) / c;
  if (nwin > 32) return set_error(ctx, ZKB_EINVAL, "window width too small for precomputation (more than 32 windows)");
  if (b->n * nwin >= (1ull << 31)) return set_error(ctx, ZKB_EINVAL, "expanded base table exceeds 2^31 points");
  const size_t sz = b->is_g2 ? 128 : 64;
  void* d_new = nullptr;
  ZKB_CUDA(ctx, cudaMalloc(&d_new, b->n * nwin * sz));
  if (b->is_g2) msm_precompute_g2(b->d_pts, d_new, b->n, nwin, c, ctx->stream);
  else msm_precompute_g1(b->d_pts, d_new, b->n, nwin, c, ctx->stream);
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) {
    cudaFree(d_new);
    return cuda_fail(ctx, e, "bases precompute");
  }
  cudaFree(b->d_pts);
  b->d_pts = d_new;
  b->precomp_c = c;
  b->precomp_nwin = nwin;
  return ZKB_OK;
}

int bases_download_nolock(zkb_ctx* ctx, const zkb_bases* b, uint64_t first, uint64_t count,
                          uint8_t* out) {
  if (!b || !out || first > b->n || count > b->n - first) return set_error(ctx, ZKB_EINVAL, "bad range");
  if (!count) return ZKB_OK;
  const size_t sz = b->is_g2 ? 128 : 64;
  ZKB_CUDA(ctx, ctx->io_b.ensure(count * sz));
  uint32_t blocks = uint32_t((count + 127) / 128);
  const char* src = reinterpret_cast<const char*>(b->d_pts) + first * sz;
  if (b->is_g2)
    bases_from_mont_kernel<Fq2><<<blocks, 128, 0, ctx->stream>>>(
        reinterpret_cast<const Affine<Fq2>*>(src), reinterpret_cast<Affine<Fq2>*>(ctx->io_b.p), count);
  else
    bases_from_mont_kernel<Fq><<<blocks, 128, 0, ctx->stream>>>(
        reinterpret_cast<const Affine<Fq>*>(src), reinterpret_cast<Affine<Fq>*>(ctx->io_b.p), count);
  count_launches(1);
  ZKB_CUDA(ctx, cudaMemcpyAsync(out, ctx->io_b.p, count * sz, cudaMemcpyDeviceToHost, ctx->stream));
  ZKB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ZKB_OK;
}

// ---------------------------------------------------------------- integer-pipe microbenchmark
// 8 independent 2-instruction carry chains per thread, ILP high enough to saturate the pipe.
__global__ void __launch_bounds__(256) imad_chain_kernel(uint32_t* sink, int iters, uint32_t seed,
                                                         long long* cycles) {
  uint32_t a = seed + threadIdx.x, b = seed * 3 + blockIdx.x;
  uint32_t r[16];
#pragma unroll
  for (int i = 0; i < 16; i++) r[i] = a + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int rep = 0; rep < 4; rep++) {
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;"
                     : "+r"(r[i]), "+r"(r[i + 1]) : "r"(a), "r"(b));
      }
    }
  }
  long long t1 = clock64();
  uint32_t x = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) x ^= r[i];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = x;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

__global__ void __launch_bounds__(256) montmul_chain_kernel(uint32_t* sink, int iters, uint32_t seed) {
  Fq a = Fq::one(), b = Fq::r2(), c = Fq::modulus(), d = Fq::one();
  a.v[0] ^= seed + threadIdx.x;
  b.v[1] ^= blockIdx.x;
  c.v[0] -= 7 + threadIdx.x;
  d.v[2] ^= seed;
  for (int it = 0; it < iters; it++) {
    a = a * b;
    c = c * d;
    b = b * a;
    d = d * c;
  }
  Fq s = a + b + c + d;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s.v[0] ^ s.v[7];
}

}  // namespace zkb

// ================================================================ extern "C"
extern "C" {

const char* zkb_version(void) { return "zkb 0.1.0 (sm_100a)"; }

int zkb_ctx_create(int device, zkb_ctx** out) {
  if (!out) return ZKB_EINVAL;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
    cudaGetLastError();
    return ZKB_ENOGPU;
  }
  if (device < 0 || device >= count) return ZKB_EINVAL;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return ZKB_ECUDA;
  if (prop.major != 10) return ZKB_ENOGPU;   // the only code in this library is sm_100a SASS
  if (cudaSetDevice(device) != cudaSuccess) return ZKB_ECUDA;
  zkb_ctx* ctx = new (std::nothrow) zkb_ctx();
  if (!ctx) return ZKB_ENOMEM;
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete ctx;
    return ZKB_ECUDA;
  }
  *out = ctx;
  return ZKB_OK;
}

void zkb_ctx_destroy(zkb_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  comm_destroy(ctx);
  for (auto& kv : ctx->dist_tables) {
    kv.second->release();
    delete kv.second;
  }
  ctx->dist_small.release();
  for (auto& kv : ctx->domains) {
    NttDomain* d = kv.second;
    d->tw_lo.release(); d->tw_hi.release(); d->tw_hi_ninv.release();
    for (auto& f : d->full) f.release();
    delete d;
  }
  for (auto& kv : ctx->cosets) {
    CosetTables* t = kv.second;
    t->fwd.release(); t->inv.release(); t->inv_z.release(); t->ninv.release();
    delete t;
  }
  ctx->stage_fwd.release(); ctx->stage_inv.release(); ctx->ntt_tmp.release();
  ctx->io_a.release(); ctx->io_b.release(); ctx->io_c.release();
  MsmScratch& s = ctx->msm;
  DevBuf* all[] = {&s.keys_in, &s.vals_in, &s.keys_out, &s.vals_out, &s.sort_temp, &s.m_valid,
                   &s.buckets, &s.part1_keys, &s.part1_pts, &s.part2_keys, &s.part2_pts,
                   &s.group_sums, &s.window_sums, &s.result_xyzz, &s.result_affine};
  for (DevBuf* b : all) b->release();
  for (int i = 0; i < 6; i++)
    if (ctx->msm_ev[i]) cudaEventDestroy(ctx->msm_ev[i]);
  for (int i = 0; i < 16; i++)
    if (ctx->copy_ev[i]) cudaEventDestroy(ctx->copy_ev[i]);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  for (int i = 0; i < 5; i++)
    if (ctx->side[i]) cudaStreamDestroy(ctx->side[i]);
  for (int i = 0; i < 6; i++)
    if (ctx->side_ev[i]) cudaEventDestroy(ctx->side_ev[i]);
  ctx->msm.multi_part.release();
  ctx->msm.group_sums2.release();
  ctx->msm.window_sums2.release();
  ctx->msm.buckets2.release();
  ctx->msm.flag.release();
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* zkb_last_error(const zkb_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
void* zkb_ctx_stream(zkb_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

// page-locked host staging (witness / scalar buffers the host fills and the library copies from)
int zkb_host_alloc(uint64_t bytes, void** out) {
  if (!out) return ZKB_EINVAL;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
    cudaGetLastError();
    return ZKB_ENOGPU;
  }
  cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable);
  if (e != cudaSuccess) {
    cudaGetLastError();
    *out = nullptr;
    return e == cudaErrorMemoryAllocation ? ZKB_ENOMEM : ZKB_ECUDA;
  }
  return ZKB_OK;
}
void zkb_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

int zkb_ctx_sync(zkb_ctx* ctx) {
  if (!ctx) return ZKB_EINVAL;
  ZKB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ZKB_OK;
}

int zkb_ctx_set(zkb_ctx* ctx, const char* key, int64_t value) {
  if (!ctx || !key) return ZKB_EINVAL;
  if (!std::strcmp(key, "msm_window")) {
    if (value < 0 || value > 20 || value == 1) return set_error(ctx, ZKB_EINVAL, "msm_window must be 0 or 2..20");
    ctx->msm_window = value;
    return ZKB_OK;
  }
  if (!std::strcmp(key, "msm_host_chunks")) {
    if (value < 0 || value > 16) return set_error(ctx, ZKB_EINVAL, "msm_host_chunks must be 0..16");
    ctx->msm_host_chunks = value;
    return ZKB_OK;
  }
  if (!std::strcmp(key, "pk_window")) {
    if (value != 0 && (value < 8 || value > 24)) return set_error(ctx, ZKB_EINVAL, "pk_window must be 0 or 8..24");
    ctx->pk_window = value;
    return ZKB_OK;
  }
  if (!std::strcmp(key, "assemble_on_device")) {
    ctx->assemble_on_device = value ? 1 : 0;
    return ZKB_OK;
  }
  if (!std::strcmp(key, "pk_precompute")) {
    ctx->pk_precompute = value ? 1 : 0;
    return ZKB_OK;
  }
  if (!std::strcmp(key, "comm_force_staging")) {   // before the first distributed call
    ctx->comm_force_staging = value ? 1 : 0;
    return ZKB_OK;
  }
  if (!std::strcmp(key, "msm_sort")) {        // process-wide: 0 = hand-written bucket sort, 1 = cub radix sort
    msm_set_sort_mode(value ? 1 : 0);
    return ZKB_OK;
  }
  if (!std::strcmp(key, "msm_g2_pair")) {     // process-wide comparison switch, see msm_g2.cu
    msm_set_g2_pair(int(value) & 3);
    return ZKB_OK;
  }
  if (!std::strcmp(key, "msm_profile")) {
    ctx->msm_profile = value ? 1 : 0;
    return ZKB_OK;
  }
  if (!std::strcmp(key, "ntt_full_twiddles")) {
    ctx->ntt_full_twiddles = value ? 1 : 0;
    return ZKB_OK;
  }
  return set_error(ctx, ZKB_EINVAL, std::string("unknown tunable ") + key);
}

#define ZKB_ENTER(ctx)                                         \
  if (!(ctx)) return ZKB_EINVAL;                               \
  ::zkb::BusyGuard _guard(ctx);                                \
  if (!_guard.ok) return ZKB_EBUSY;                            \
  ZKB_CUDA(ctx, cudaSetDevice((ctx)->device));

// ---- NTT
int zkb_fr_ntt_dev(zkb_ctx* ctx, const void* d_in, void* d_out, uint32_t log_n, int inverse,
                   int coset, int montgomery) {
  ZKB_ENTER(ctx);
  if (!d_in || !d_out) return set_error(ctx, ZKB_EINVAL, "null buffer");
  const void *pre = nullptr, *post = nullptr;
  if (coset) {
    CosetTables* t;
    int rc = coset_tables(ctx, log_n, &t);
    if (rc) return rc;
    if (inverse) post = t->inv.p; else pre = t->fwd.p;
  }
  return ntt_run(ctx, d_in, d_out, log_n, inverse, pre, post, 0, montgomery, montgomery);
}

// ---- building blocks of the distributed (four-step) NTT; all asynchronous on device buffers
int zkb_fr_ntt_batch_dev(zkb_ctx* ctx, const void* d_in, void* d_out, uint32_t log_size, uint64_t batch,
                         int inverse, int skip_ninv, const void* d_pre_table, const void* d_post_table) {
  ZKB_ENTER(ctx);
  if (!d_in || !d_out) return set_error(ctx, ZKB_EINVAL, "null buffer");
  return ntt_run(ctx, d_in, d_out, log_size, inverse, d_pre_table, d_post_table, skip_ninv, 1, 1, batch);
}
int zkb_fr_gather_layout_dev(zkb_ctx* ctx, const void* d_full, void* d_out, uint32_t log_n, uint32_t low_bits,
                             uint32_t world, uint32_t rank) {
  ZKB_ENTER(ctx);
  if (!d_full || !d_out || low_bits > log_n || world == 0 || rank >= world || ((1u << low_bits) % world))
    return set_error(ctx, ZKB_EINVAL, "bad layout argument");
  fr_gather_layout_launch(d_full, d_out, log_n, low_bits, world, rank, ctx->stream);
  ZKB_CUDA(ctx, cudaGetLastError());
  return ZKB_OK;
}
int zkb_fr_table_dev(zkb_ctx* ctx, void* d_out, const uint8_t base[32], const uint8_t scale[32], uint64_t rows,
                     uint64_t cols, uint64_t row0, uint64_t mul, int mode) {
  ZKB_ENTER(ctx);
  HFr b, c;
  if (!d_out || !base || !scale || !HFr::from_bytes(base, &b) || !HFr::from_bytes(scale, &c) || mode < 0 || mode > 1)
    return set_error(ctx, ZKB_EINVAL, "bad table argument");
  uint32_t bw[8], cw[8];
  hfr_to_words(b, bw);
  hfr_to_words(c, cw);
  fr_table_launch(d_out, bw, cw, rows, cols, row0, mul, mode, ctx->stream);
  ZKB_CUDA(ctx, cudaGetLastError());
  return ZKB_OK;
}
int zkb_fr_mul_sub_dev(zkb_ctx* ctx, void* d_a, const void* d_b, const void* d_c, uint64_t n) {
  ZKB_ENTER(ctx);
  if (!d_a || !d_b || !d_c) return set_error(ctx, ZKB_EINVAL, "null buffer");
  fr_mul_sub_launch(d_a, d_b, d_c, n, ctx->stream);
  ZKB_CUDA(ctx, cudaGetLastError());
  return ZKB_OK;
}
int zkb_fr_convert_dev(zkb_ctx* ctx, void* d_data, uint64_t n, int to_montgomery) {
  ZKB_ENTER(ctx);
  if (!d_data) return set_error(ctx, ZKB_EINVAL, "null buffer");
  fr_convert_launch(d_data, n, to_montgomery, ctx->stream);
  ZKB_CUDA(ctx, cudaGetLastError());
  return ZKB_OK;
}
void zkb_fr_root_of_unity(uint32_t log_n, int inverse, uint8_t out[32]) {
  uint32_t w[8];
  host_fr_root(log_n, inverse, w);
  HFr a;
  std::memcpy(a.l, w, 32);
  a.to_bytes(out);
}

int zkb_fr_ntt(zkb_ctx* ctx, uint8_t* data, uint32_t log_n, int inverse, int coset) {
  if (!ctx) return ZKB_EINVAL;
  if (!data) return set_error(ctx, ZKB_EINVAL, "null buffer");
  if (log_n < 1 || log_n > 28) return set_error(ctx, ZKB_EINVAL, "log_n must be in [1, 28]");
  uint64_t n = 1ull << log_n;
  {
    ZKB_ENTER(ctx);
    ZKB_CUDA(ctx, ctx->io_a.ensure(n * 32));
    ZKB_CUDA(ctx, ctx->msm.flag.ensure(16));
    ZKB_CUDA(ctx, cudaMemsetAsync(ctx->msm.flag.p, 0, 16, ctx->stream));
    ZKB_CUDA(ctx, cudaMemcpyAsync(ctx->io_a.p, data, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  }
  // out of place on the device so that a rejected input leaves the caller's buffer untouched
  int rc;
  {
    ZKB_ENTER(ctx);
    ZKB_CUDA(ctx, ctx->io_b.ensure(n * 32));
  }
  rc = zkb_fr_ntt_dev(ctx, ctx->io_a.p, ctx->io_b.p, log_n, inverse, coset, 0);
  if (rc) return rc;
  ZKB_ENTER(ctx);
  uint32_t flag = 0;
  ZKB_CUDA(ctx, cudaMemcpyAsync(&flag, ctx->msm.flag.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
  ZKB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (flag & 1) {
    ZKB_CUDA(ctx, cudaMemsetAsync(ctx->msm.flag.p, 0, 16, ctx->stream));
    return set_error(ctx, ZKB_EINVAL, "non-canonical Fr element");
  }
  ZKB_CUDA(ctx, cudaMemcpyAsync(data, ctx->io_b.p, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
  ZKB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ZKB_OK;
}

// ---- bases
static int bases_alloc(zkb_ctx* ctx, uint64_t n, int is_g2, zkb_bases** out) {
  zkb_bases* b = new (std::nothrow) zkb_bases();
  if (!b) return set_error(ctx, ZKB_ENOMEM, "out of host memory");
  b->device = ctx->device;
  b->is_g2 = is_g2;
  b->n = n;
  cudaError_t e = cudaMalloc(&b->d_pts, (n ? n : 1) * (is_g2 ? 128 : 64));
  if (e != cudaSuccess) {
    delete b;
    return cuda_fail(ctx, e, "bases allocation");
  }
  *out = b;
  return ZKB_OK;
}

static int bases_upload(zkb_ctx* ctx, const uint8_t* affine, uint64_t n, int is_g2, zkb_bases** out) {
  ZKB_ENTER(ctx);
  if (!out || (!affine && n)) return set_error(ctx, ZKB_EINVAL, "null argument");
  const size_t sz = is_g2 ? 128 : 64;
  for (uint64_t i = 0; i < n * (sz / 32); i++) {
    uint64_t w[4];
    std::memcpy(w, affine + 32 * i, 32);
    if (!host::HFq::lt_mod(w)) return set_error(ctx, ZKB_EINVAL, "non-canonical Fq coordinate");
  }
  zkb_bases* b;
  int rc = bases_alloc(ctx, n, is_g2, &b);
  if (rc) return rc;
  cudaError_t e = cudaMemcpyAsync(b->d_pts, affine, n * sz, cudaMemcpyHostToDevice, ctx->stream);
  if (e != cudaSuccess) {
    zkb_bases_free(b);
    return cuda_fail(ctx, e, "bases upload");
  }
  if ((rc = bases_to_mont_checked(ctx, b))) {
    zkb_bases_free(b);
    return rc;
  }
  *out = b;
  return ZKB_OK;
}

int zkb_bases_upload_g1(zkb_ctx* ctx, const uint8_t* affine, uint64_t n, zkb_bases** out) {
  return bases_upload(ctx, affine, n, 0, out);
}
int zkb_bases_upload_g2(zkb_ctx* ctx, const uint8_t* affine, uint64_t n, zkb_bases** out) {
  return bases_upload(ctx, affine, n, 1, out);
}

static int bases_synthetic(zkb_ctx* ctx, uint64_t n, uint64_t seed, int is_g2, zkb_bases** out) {
  ZKB_ENTER(ctx);
  if (!out) return set_error(ctx, ZKB_EINVAL, "null argument");
  std::vector<uint64_t> k;
  try {
    k.resize(4 * n);
  } catch (...) {
    return set_error(ctx, ZKB_ENOMEM, "out of host memory");
  }
  host::Xoshiro256ss rng(seed);
  for (uint64_t i = 0; i < n; i++) rng.fr_canonical(&k[4 * i]);
  zkb_bases* b;
  int rc = bases_alloc(ctx, n, is_g2, &b);
  if (rc) return rc;
  cudaError_t e = ctx->io_a.ensure(n * 32 + 32);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(ctx->io_a.p, k.data(), n * 32, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) {
    rc = bases_from_scalars(ctx, ctx->io_a.p, n, is_g2, b->d_pts);
    e = cudaStreamSynchronize(ctx->stream);
  }
  if (e != cudaSuccess || rc) {
    zkb_bases_free(b);
    return rc ? rc : cuda_fail(ctx, e, "synthetic bases");
  }
  *out = b;
  return ZKB_OK;
}
int zkb_bases_synthetic_g1(zkb_ctx* ctx, uint64_t n, uint64_t seed, zkb_bases** out) {
  return bases_synthetic(ctx, n, seed, 0, out);
}
int zkb_bases_synthetic_g2(zkb_ctx* ctx, uint64_t n, uint64_t seed, zkb_bases** out) {
  return bases_synthetic(ctx, n, seed, 1, out);
}

int zkb_bases_download(zkb_ctx* ctx, const zkb_bases* b, uint64_t first, uint64_t count,
                       uint8_t* out) {
  ZKB_ENTER(ctx);
  return bases_download_nolock(ctx, b, first, count, out);
}

uint64_t zkb_bases_len(const zkb_bases* b) { return b ? b->n : 0; }

int zkb_bases_precompute(zkb_ctx* ctx, zkb_bases* b, int window_bits) {
  ZKB_ENTER(ctx);
  return bases_precompute_nolock(ctx, b, window_bits);
}
int zkb_bases_window(const zkb_bases* b) { return b ? b->precomp_c : 0; }

void zkb_bases_free(zkb_bases* b) {
  if (!b) return;
  int cur = 0;
  cudaGetDevice(&cur);
  cudaSetDevice(b->device);
  if (b->d_pts) cudaFree(b->d_pts);
  cudaSetDevice(cur);
  delete b;
}

// ---- MSM
static int msm_dev(zkb_ctx* ctx, const zkb_bases* b, uint64_t first, const void* d_scalars,
                   uint64_t n, void* d_out, int want_g2) {
  ZKB_ENTER(ctx);
  if (!b || !d_out || (!d_scalars && n)) return set_error(ctx, ZKB_EINVAL, "null argument");
  if (b->is_g2 != want_g2) return set_error(ctx, ZKB_EINVAL, "bases are of the other group");
  return msm_exec(ctx, b, first, d_scalars, n, d_out);
}
int zkb_msm_g1_dev(zkb_ctx* ctx, const zkb_bases* b, uint64_t first, const void* d_scalars,
                   uint64_t n, void* d_out) {
  return msm_dev(ctx, b, first, d_scalars, n, d_out, 0);
}
int zkb_msm_g2_dev(zkb_ctx* ctx, const zkb_bases* b, uint64_t first, const void* d_scalars,
                   uint64_t n, void* d_out) {
  return msm_dev(ctx, b, first, d_scalars, n, d_out, 1);
}

static int msm_host(zkb_ctx* ctx, const zkb_bases* b, uint64_t first, const uint8_t* scalars,
                    uint64_t n, uint8_t* out, int want_g2) {
  ZKB_ENTER(ctx);
  if (!b || !out || (!scalars && n)) return set_error(ctx, ZKB_EINVAL, "null argument");
  if (b->is_g2 != want_g2) return set_error(ctx, ZKB_EINVAL, "bases are of the other group");
  const size_t aff = want_g2 ? 128 : 64;
  ZKB_CUDA(ctx, ctx->io_c.ensure(256));
  int rc = msm_exec_host(ctx, b, first, scalars, n, ctx->io_c.p);
  if (rc) return rc;
  // result and the canonicity flag raised by the digit kernel come back together
  uint32_t flag = 0;
  ZKB_CUDA(ctx, cudaMemcpyAsync(out, ctx->io_c.p, aff, cudaMemcpyDeviceToHost, ctx->stream));
  if (ctx->msm.flag.p)
    ZKB_CUDA(ctx, cudaMemcpyAsync(&flag, ctx->msm.flag.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
  ZKB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (flag & 1) {
    ZKB_CUDA(ctx, cudaMemsetAsync(ctx->msm.flag.p, 0, 16, ctx->stream));
    std::memset(out, 0, aff);
    return set_error(ctx, ZKB_EINVAL, "non-canonical scalar");
  }
  return ZKB_OK;
}
int zkb_msm_g1(zkb_ctx* ctx, const zkb_bases* b, uint64_t first, const uint8_t* scalars,
               uint64_t n, uint8_t out[64]) {
  return msm_host(ctx, b, first, scalars, n, out, 0);
}
int zkb_msm_g2(zkb_ctx* ctx, const zkb_bases* b, uint64_t first, const uint8_t* scalars,
               uint64_t n, uint8_t out[128]) {
  return msm_host(ctx, b, first, scalars, n, out, 1);
}

static int point_sum(zkb_ctx* ctx, const uint8_t* points, uint32_t n, uint8_t* out, int is_g2) {
  ZKB_ENTER(ctx);
  if (!out || (!points && n)) return set_error(ctx, ZKB_EINVAL, "null argument");
  const size_t aff = is_g2 ? 128 : 64;
  ZKB_CUDA(ctx, ctx->io_b.ensure(size_t(n) * aff + aff));
  ZKB_CUDA(ctx, ctx->io_c.ensure(256));
  if (n) ZKB_CUDA(ctx, cudaMemcpyAsync(ctx->io_b.p, points, size_t(n) * aff, cudaMemcpyHostToDevice, ctx->stream));
  if (is_g2) msm_sum_affine_g2(ctx->io_b.p, n, ctx->io_c.p, ctx->stream);
  else msm_sum_affine_g1(ctx->io_b.p, n, ctx->io_c.p, ctx->stream);
  ZKB_CUDA(ctx, cudaMemcpyAsync(out, ctx->io_c.p, aff, cudaMemcpyDeviceToHost, ctx->stream));
  ZKB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ZKB_OK;
}
int zkb_g1_sum(zkb_ctx* ctx, const uint8_t* points, uint32_t n, uint8_t out[64]) {
  return point_sum(ctx, points, n, out, 0);
}
int zkb_g2_sum(zkb_ctx* ctx, const uint8_t* points, uint32_t n, uint8_t out[128]) {
  return point_sum(ctx, points, n, out, 1);
}

static int point_sum_dev(zkb_ctx* ctx, const void* d_points, uint32_t n, void* d_out, int is_g2) {
  ZKB_ENTER(ctx);
  if (!d_out || (!d_points && n)) return set_error(ctx, ZKB_EINVAL, "null argument");
  if (is_g2) msm_sum_affine_g2(d_points, n, d_out, ctx->stream);
  else msm_sum_affine_g1(d_points, n, d_out, ctx->stream);
  ZKB_CUDA(ctx, cudaGetLastError());
  return ZKB_OK;
}
int zkb_g1_sum_dev(zkb_ctx* ctx, const void* d_points, uint32_t n, void* d_out) {
  return point_sum_dev(ctx, d_points, n, d_out, 0);
}
int zkb_g2_sum_dev(zkb_ctx* ctx, const void* d_points, uint32_t n, void* d_out) {
  return point_sum_dev(ctx, d_points, n, d_out, 1);
}

uint64_t zkb_ctx_launches(const zkb_ctx* ctx) { return ctx ? launches_so_far() : 0; }

int zkb_msm_last_timings(zkb_ctx* ctx, float ms_out[5]) {
  ZKB_ENTER(ctx);
  if (!ms_out) return set_error(ctx, ZKB_EINVAL, "null argument");
  if (!ctx->msm_ev_valid) return set_error(ctx, ZKB_EINVAL, "no profiled MSM (set msm_profile = 1 first)");
  ZKB_CUDA(ctx, cudaEventSynchronize(ctx->msm_ev[5]));
  for (int i = 0; i < 5; i++) ZKB_CUDA(ctx, cudaEventElapsedTime(&ms_out[i], ctx->msm_ev[i], ctx->msm_ev[i + 1]));
  return ZKB_OK;
}

// ---- test hooks
int zkb_field_batch_op(zkb_ctx* ctx, int field, int op, const uint8_t* a, const uint8_t* b,
                       uint8_t* out, uint64_t n) {
  ZKB_ENTER(ctx);
  if (!a || !out || op < 0 || op > 9 || field < 0 || field > 1) return set_error(ctx, ZKB_EINVAL, "bad argument");
  if (op <= 2 && !b) return set_error(ctx, ZKB_EINVAL, "binary op needs b");
  if (!n) return ZKB_OK;
  ZKB_CUDA(ctx, ctx->io_a.ensure(n * 32));
  ZKB_CUDA(ctx, ctx->io_b.ensure(n * 32));
  ZKB_CUDA(ctx, ctx->io_c.ensure(n * 32 > 256 ? n * 32 : 256));
  ZKB_CUDA(ctx, cudaMemcpyAsync(ctx->io_a.p, a, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  if (b) ZKB_CUDA(ctx, cudaMemcpyAsync(ctx->io_b.p, b, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  uint32_t blocks = uint32_t((n + 127) / 128);
  if (field == 0)
    field_batch_kernel<Fq><<<blocks, 128, 0, ctx->stream>>>(op, (const Fq*)ctx->io_a.p, b ? (const Fq*)ctx->io_b.p : nullptr, (Fq*)ctx->io_c.p, n);
  else
    field_batch_kernel<Fr><<<blocks, 128, 0, ctx->stream>>>(op, (const Fr*)ctx->io_a.p, b ? (const Fr*)ctx->io_b.p : nullptr, (Fr*)ctx->io_c.p, n);
    count_launches(1);
  ZKB_CUDA(ctx, cudaMemcpyAsync(out, ctx->io_c.p, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
  ZKB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ZKB_OK;
}

int zkb_curve_batch_op(zkb_ctx* ctx, int group, int op, const uint8_t* a, const uint8_t* b,
                       uint8_t* out, uint64_t n) {
  ZKB_ENTER(ctx);
  if (!a || !out || op < 0 || op > 2 || group < 0 || group > 1) return set_error(ctx, ZKB_EINVAL, "bad argument");
  if (op != 1 && !b) return set_error(ctx, ZKB_EINVAL, "op needs b");
  if (!n) return ZKB_OK;
  const size_t aff = group ? 128 : 64;
  const size_t bsz = op == 2 ? 32 : aff;
  ZKB_CUDA(ctx, ctx->io_a.ensure(n * aff));
  ZKB_CUDA(ctx, ctx->io_b.ensure(n * bsz));
  ZKB_CUDA(ctx, ctx->io_c.ensure(n * aff > 256 ? n * aff : 256));
  ZKB_CUDA(ctx, cudaMemcpyAsync(ctx->io_a.p, a, n * aff, cudaMemcpyHostToDevice, ctx->stream));
  if (b) ZKB_CUDA(ctx, cudaMemcpyAsync(ctx->io_b.p, b, n * bsz, cudaMemcpyHostToDevice, ctx->stream));
  uint32_t blocks = uint32_t((n + 63) / 64);
  if (group == 0)
    curve_batch_kernel<Fq><<<blocks, 64, 0, ctx->stream>>>(op, (const Affine<Fq>*)ctx->io_a.p, ctx->io_b.p, (Affine<Fq>*)ctx->io_c.p, n);
  else
    curve_batch_kernel<Fq2><<<blocks, 64, 0, ctx->stream>>>(op, (const Affine<Fq2>*)ctx->io_a.p, ctx->io_b.p, (Affine<Fq2>*)ctx->io_c.p, n);
    count_launches(1);
  ZKB_CUDA(ctx, cudaMemcpyAsync(out, ctx->io_c.p, n * aff, cudaMemcpyDeviceToHost, ctx->stream));
  ZKB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ZKB_OK;
}

int zkb_bench_int_pipe(zkb_ctx* ctx, int iters, double results[3]) {
  ZKB_ENTER(ctx);
  if (!results || iters <= 0) return set_error(ctx, ZKB_EINVAL, "bad argument");
  const int blocks = ctx->sm_count * 8, threads = 256;
  ZKB_CUDA(ctx, ctx->io_a.ensure(size_t(blocks) * threads * 4 + 64));
  ZKB_CUDA(ctx, ctx->io_c.ensure(256));
  cudaEvent_t e0, e1;
  ZKB_CUDA(ctx, cudaEventCreate(&e0));
  ZKB_CUDA(ctx, cudaEventCreate(&e1));
  float ms = 0;
  // IMAD.WIDE carry chains: 32 wide-IMADs per inner iteration per thread
  imad_chain_kernel<<<blocks, threads, 0, ctx->stream>>>((uint32_t*)ctx->io_a.p, iters / 8 + 1, 1, (long long*)ctx->io_c.p);
  count_launches(1);
  cudaEventRecord(e0, ctx->stream);
  imad_chain_kernel<<<blocks, threads, 0, ctx->stream>>>((uint32_t*)ctx->io_a.p, iters, 12345, (long long*)ctx->io_c.p);
  count_launches(1);
  cudaEventRecord(e1, ctx->stream);
  ZKB_CUDA(ctx, cudaEventSynchronize(e1));
  cudaEventElapsedTime(&ms, e0, e1);
  long long cycles = 0;
  ZKB_CUDA(ctx, cudaMemcpy(&cycles, ctx->io_c.p, 8, cudaMemcpyDeviceToHost));
  results[0] = double(blocks) * threads * double(iters) * 32.0 / (ms * 1e-3);
  results[2] = double(cycles) / (ms * 1e-3) / 1e6;
  // Montgomery multiplications: 4 per iteration per thread
  montmul_chain_kernel<<<blocks, threads, 0, ctx->stream>>>((uint32_t*)ctx->io_a.p, iters / 8 + 1, 1);
  count_launches(1);
  cudaEventRecord(e0, ctx->stream);
  montmul_chain_kernel<<<blocks, threads, 0, ctx->stream>>>((uint32_t*)ctx->io_a.p, iters, 777);
  count_launches(1);
  cudaEventRecord(e1, ctx->stream);
  ZKB_CUDA(ctx, cudaEventSynchronize(e1));
  cudaEventElapsedTime(&ms, e0, e1);
  results[1] = double(blocks) * threads * double(iters) * 4.0 / (ms * 1e-3);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  ZKB_CUDA(ctx, cudaGetLastError());
  return ZKB_OK;
}

}  // extern "C"
