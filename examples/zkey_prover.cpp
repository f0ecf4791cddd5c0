// zkey_prover.cpp — the C ABI (include/zkb.h) as a file-to-file prover with the command line of the
// snarkjs / rapidsnark provers that sit behind eigen-zeth's prover service today:
//
//   zkey_prover <circuit.zkey> <witness.wtns> <proof.json> <public.json>
//
// Reads a snarkjs Groth16 proving key and a circom witness (both parsed by libzkb.so), loads the key
// with h_basis = 1, proves on the GPU with fresh blinding scalars, checks the pairing equation under
// the key's own verifying key and writes the two files in the layout eigen-zeth stores and parses (reference
// proof/proof.json:1-22, proof/public_input.json:1-3, src/settlement/ethereum/mod.rs:445-481).
// Exit codes: 0 proof written, 1 bad input / prover error, 2 no GPU (there is no CPU fallback).
// Both files are parsed BEFORE the GPU is touched, so a malformed key fails the same way anywhere.
//
//   g++ -std=c++17 -Iinclude examples/zkey_prover.cpp -Leigen-zeth_b200 -lzkb -Wl,-rpath,$PWD/eigen-zeth_b200 -o zkey_prover
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "zkb.h"

static bool read_file(const char* path, std::vector<uint8_t>* out) {
  FILE* f = fopen(path, "rb");
  if (!f) return false;
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  if (n < 0) { fclose(f); return false; }
  out->resize(size_t(n));
  bool ok = n == 0 || fread(out->data(), 1, size_t(n), f) == size_t(n);
  fclose(f);
  return ok;
}

// 32-byte little-endian integer -> decimal string (schoolbook division by 10^9)
static std::string to_decimal(const uint8_t* le) {
  uint32_t w[8];
  for (int i = 0; i < 8; i++) w[i] = uint32_t(le[4 * i]) | uint32_t(le[4 * i + 1]) << 8 | uint32_t(le[4 * i + 2]) << 16 | uint32_t(le[4 * i + 3]) << 24;
  std::string out;
  for (;;) {
    uint64_t rem = 0;
    bool nz = false;
    for (int i = 7; i >= 0; i--) {
      uint64_t cur = (rem << 32) | w[i];
      w[i] = uint32_t(cur / 1000000000u);
      rem = cur % 1000000000u;
      nz = nz || w[i];
    }
    char buf[16];
    snprintf(buf, sizeof buf, nz ? "%09u" : "%u", unsigned(rem));
    out = buf + out;
    if (!nz) break;
  }
  return out;
}

// a blinding scalar: 31 random bytes (< 2^248 < r, so canonical without a reduction)
static bool random_scalar(uint8_t out[32]) {
  FILE* f = fopen("/dev/urandom", "rb");
  if (!f) return false;
  std::memset(out, 0, 32);
  bool ok = fread(out, 1, 31, f) == 31;
  fclose(f);
  return ok;
}

int main(int argc, char** argv) {
  if (argc != 5) {
    fprintf(stderr, "usage: %s <circuit.zkey> <witness.wtns> <proof.json> <public.json>\n", argv[0]);
    return 1;
  }
  std::vector<uint8_t> zfile, wfile;
  if (!read_file(argv[1], &zfile) || !read_file(argv[2], &wfile)) {
    fprintf(stderr, "cannot read %s or %s\n", argv[1], argv[2]);
    return 1;
  }
  zkb_zkey_info zi{};
  if (zkb_zkey_inspect(zfile.data(), zfile.size(), &zi) != ZKB_OK) {
    fprintf(stderr, "%s: not a BN254 Groth16 .zkey\n", argv[1]);
    return 1;
  }
  uint32_t n_wit = 0;
  if (zkb_wtns_inspect(wfile.data(), wfile.size(), &n_wit) != ZKB_OK || n_wit != zi.n_vars) {
    fprintf(stderr, "%s: not a BN254 .wtns for this key (%u values, the key has %u variables)\n", argv[2], n_wit, zi.n_vars);
    return 1;
  }
  const uint64_t m = zi.n_vars, np = uint64_t(zi.n_public) + 1, n = 1ull << zi.log_n;
  std::vector<uint32_t> row_ptr[2], col[2];
  std::vector<uint8_t> coeff[2];
  for (int k = 0; k < 2; k++) {
    row_ptr[k].resize(zi.n_rows + 1);
    col[k].resize(zi.nnz[k] ? zi.nnz[k] : 1);
    coeff[k].resize(zi.nnz[k] ? 32 * zi.nnz[k] : 1);
  }
  std::vector<uint8_t> fixed(64 * 3 + 128 * 3), ic(64 * np), a(64 * m), b1(64 * m), b2(128 * m), c(64 * (m - np) + 1), h(64 * n),
      witness(32 * m);
  zkb_zkey_data zd{};
  for (int k = 0; k < 2; k++) {
    zd.row_ptr[k] = row_ptr[k].data();
    zd.col[k] = col[k].data();
    zd.coeff[k] = coeff[k].data();
  }
  zd.alpha_g1 = fixed.data();
  zd.beta_g1 = fixed.data() + 64;
  zd.delta_g1 = fixed.data() + 128;
  zd.beta_g2 = fixed.data() + 192;
  zd.gamma_g2 = fixed.data() + 320;
  zd.delta_g2 = fixed.data() + 448;
  zd.ic = ic.data();
  zd.a_query = a.data();
  zd.b_g1_query = b1.data();
  zd.b_g2_query = b2.data();
  zd.c_query = c.data();
  zd.h_query = h.data();
  if (zkb_zkey_read(zfile.data(), zfile.size(), &zd) != ZKB_OK) {
    fprintf(stderr, "%s: a coordinate or coefficient is not reduced\n", argv[1]);
    return 1;
  }
  if (zkb_wtns_read(wfile.data(), wfile.size(), witness.data()) != ZKB_OK) {
    fprintf(stderr, "%s: a value is not reduced or wire 0 is not 1\n", argv[2]);
    return 1;
  }
  uint8_t r[32], s[32];
  if (!random_scalar(r) || !random_scalar(s)) {
    fprintf(stderr, "no randomness for the blinding scalars\n");
    return 1;
  }

  zkb_ctx* ctx = nullptr;
  int rc = zkb_ctx_create(0, &ctx);
  if (rc) {
    fprintf(stderr, "zkb_ctx_create failed (%d)%s\n", rc, rc == ZKB_ENOGPU ? ": no sm_100 GPU, and there is no CPU fallback" : "");
    return 2;
  }
  zkb_pk_desc d{};
  d.log_n = zi.log_n;
  d.n_public = zi.n_public;
  d.n_vars = m;
  d.n_constraints = zi.n_rows;
  for (int k = 0; k < 2; k++) {
    d.row_ptr[k] = row_ptr[k].data();
    d.col[k] = col[k].data();
    d.coeff[k] = coeff[k].data();
  }
  d.a_query = a.data();
  d.b_g1_query = b1.data();
  d.b_g2_query = b2.data();
  d.l_query = c.data();
  d.h_query = h.data();
  d.alpha_g1 = zd.alpha_g1;
  d.beta_g1 = zd.beta_g1;
  d.delta_g1 = zd.delta_g1;
  d.beta_g2 = zd.beta_g2;
  d.delta_g2 = zd.delta_g2;
  d.h_basis = 1;
  zkb_pk* pk = nullptr;
  if ((rc = zkb_pk_load(ctx, &d, &pk))) {
    fprintf(stderr, "key: %s\n", zkb_last_error(ctx));
    return 1;
  }
  uint8_t proof[ZKB_PROOF_BYTES];
  zkb_prove_timings t{};
  if ((rc = zkb_groth16_prove(ctx, pk, witness.data(), r, s, proof, nullptr, nullptr, &t))) {
    fprintf(stderr, "prove: %s\n", zkb_last_error(ctx));
    return 1;
  }
  uint8_t vk_fixed[448];
  std::memcpy(vk_fixed, zd.alpha_g1, 64);
  std::memcpy(vk_fixed + 64, zd.beta_g2, 128);
  std::memcpy(vk_fixed + 192, zd.gamma_g2, 128);
  std::memcpy(vk_fixed + 320, zd.delta_g2, 128);
  const bool ok = zkb_groth16_verify_ex(vk_fixed, ic.data(), zi.n_public, proof, witness.data() + 32) == ZKB_OK;
  fprintf(stderr, "2^%u domain, %u variables, %u public: %.2f ms on the device; pairing check %s\n", zi.log_n, zi.n_vars,
          zi.n_public, t.ms_total, ok ? "passed" : "FAILED (witness does not satisfy the circuit?)");
  if (!ok) return 1;
  FILE* fp = fopen(argv[3], "w");
  FILE* fi = fopen(argv[4], "w");
  if (!fp || !fi) {
    fprintf(stderr, "cannot write %s or %s\n", argv[3], argv[4]);
    return 1;
  }
  const uint8_t* p = proof;
  fprintf(fp, "{\n    \"pi_a\": {\n        \"x\": \"%s\",\n        \"y\": \"%s\"\n    },\n", to_decimal(p).c_str(), to_decimal(p + 32).c_str());
  fprintf(fp, "    \"pi_b\": {\n        \"x\": [\n            \"%s\",\n            \"%s\"\n        ],\n        \"y\": [\n            \"%s\",\n            \"%s\"\n        ]\n    },\n",
          to_decimal(p + 64).c_str(), to_decimal(p + 96).c_str(), to_decimal(p + 128).c_str(), to_decimal(p + 160).c_str());
  fprintf(fp, "    \"pi_c\": {\n        \"x\": \"%s\",\n        \"y\": \"%s\"\n    },\n    \"protocol\": \"groth16\",\n    \"curve\": \"BN128\"\n}\n",
          to_decimal(p + 192).c_str(), to_decimal(p + 224).c_str());
  fprintf(fi, "[\n");
  for (uint32_t i = 0; i < zi.n_public; i++)
    fprintf(fi, "    \"%s\"%s\n", to_decimal(witness.data() + 32 * (i + 1)).c_str(), i + 1 < zi.n_public ? "," : "");
  fprintf(fi, "]\n");
  fclose(fp);
  fclose(fi);
  zkb_pk_free(pk);
  zkb_ctx_destroy(ctx);
  return 0;
}
