// prove_cli.cpp — the C ABI (include/zkb.h) from compiled host code: build a synthetic proving key,
// prove, pairing-check and print proof.json / public_input.json in the layout eigen-zeth parses
// (reference proof/proof.json:1-22, src/settlement/ethereum/mod.rs:445-481).
//
//   g++ -std=c++17 -Iinclude examples/prove_cli.cpp -Leigen-zeth_b200 -lzkb -Wl,-rpath,$PWD/eigen-zeth_b200 -o prove_cli
//   ./prove_cli 16        # 2^16-constraint circuit
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "zkb.h"

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

int main(int argc, char** argv) {
  uint32_t log_n = argc > 1 ? uint32_t(atoi(argv[1])) : 12;
  zkb_ctx* ctx = nullptr;
  int rc = zkb_ctx_create(0, &ctx);
  if (rc) {
    fprintf(stderr, "zkb_ctx_create failed (%d)%s\n", rc, rc == ZKB_ENOGPU ? ": no sm_100 GPU, and there is no CPU fallback" : "");
    return 2;
  }
  zkb_pk* pk = nullptr;
  uint8_t vk[ZKB_VK_BYTES], proof[ZKB_PROOF_BYTES], w0[32] = {5}, w1[32] = {7}, r[32] = {11}, s[32] = {13};
  if ((rc = zkb_pk_synthetic(ctx, log_n, 0x5EED0004, &pk, vk))) {
    fprintf(stderr, "setup: %s\n", zkb_last_error(ctx));
    return 1;
  }
  std::vector<uint8_t> z(size_t(32) << log_n);
  zkb_synthetic_witness(log_n, w0, w1, z.data());
  zkb_prove_timings t{};
  if ((rc = zkb_groth16_prove(ctx, pk, z.data(), r, s, proof, nullptr, nullptr, &t))) {
    fprintf(stderr, "prove: %s\n", zkb_last_error(ctx));
    return 1;
  }
  int ok = zkb_groth16_verify(vk, proof, z.data() + 32);
  fprintf(stderr, "2^%u constraints: %.2f ms on the device, pairing check %s\n", log_n, t.ms_total, ok == ZKB_OK ? "passed" : "FAILED");
  const uint8_t* p = proof;
  printf("{\n    \"pi_a\": {\n        \"x\": \"%s\",\n        \"y\": \"%s\"\n    },\n", to_decimal(p).c_str(), to_decimal(p + 32).c_str());
  printf("    \"pi_b\": {\n        \"x\": [\n            \"%s\",\n            \"%s\"\n        ],\n        \"y\": [\n            \"%s\",\n            \"%s\"\n        ]\n    },\n",
         to_decimal(p + 64).c_str(), to_decimal(p + 96).c_str(), to_decimal(p + 128).c_str(), to_decimal(p + 160).c_str());
  printf("    \"pi_c\": {\n        \"x\": \"%s\",\n        \"y\": \"%s\"\n    },\n    \"protocol\": \"groth16\",\n    \"curve\": \"BN128\"\n}\n",
         to_decimal(p + 192).c_str(), to_decimal(p + 224).c_str());
  printf("[\n    \"%s\"\n]\n", to_decimal(z.data() + 32).c_str());
  zkb_pk_free(pk);
  zkb_ctx_destroy(ctx);
  return ok == ZKB_OK ? 0 : 1;
}
