// Readers for the iden3 binary circuit formats that feed zkb_pk_load / zkb_groth16_prove with an
// externally compiled circuit (the reference holds no circuit code — its prover is
// an external service, src/prover/provider.rs:606 — so the format follows the published iden3
// "binfile" layout that circom / snarkjs write):
//
//   file      := magic[4] version:u32 n_sections:u32 section*
//   section   := type:u32 size:u64 payload[size]                 (any order, unknown types skipped)
//   .r1cs  1  := n8:u32 prime[n8] n_wires:u32 n_pub_out:u32 n_pub_in:u32 n_prv_in:u32 n_labels:u64
//                n_constraints:u32
//          2  := per constraint, for A then B then C:  n_terms:u32 (wire:u32 coeff[n8])*
//          3  := wire -> label map (n_wires x u64), not needed by the prover
//   .wtns  1  := n8:u32 prime[n8] n_witness:u32
//          2  := n_witness x value[n8]
//
//   .zkey  1  := protocol:u32 (1 = groth16)
//          2  := n8q:u32 q[n8q] n8r:u32 r[n8r] n_vars:u32 n_public:u32 domain_size:u32
//                alpha_g1 beta_g1 beta_g2 gamma_g2 delta_g1 delta_g2
//          3  := IC (n_public + 1 G1)      4 := count:u32 (matrix:u32 row:u32 signal:u32 value[n8r])*
//          5..9 := A (n_vars G1), B1 (n_vars G1), B2 (n_vars G2), C (n_vars - n_public - 1 G1),
//                H (domain_size G1);   10 := contributions (not read)
//          points: affine coordinates as Montgomery-form integers (G2: c0 then c1), all-zero =
//          infinity; coefficient values: coefficient * 2^512 mod r.  The test tree holds a writer
//          for this layout; it is not pinned against a snarkjs-written file (none in this image).
//
// All integers little-endian; .r1cs / .wtns coefficients and values are canonical (non-Montgomery) integers.
// Wire 0 is the constant 1, followed by public outputs, public inputs, private inputs — the
// z = (1, public..., private...) order zkb_pk_desc uses, so n_public = n_pub_out + n_pub_in.
// Host code only; no allocation (two-pass: *_inspect sizes the caller's buffers, *_read fills them).
#include <cstdint>
#include <cstring>

#include "../../include/zkb.h"
#include "host_ff.hpp"

namespace {

// BN254 scalar field modulus r, little-endian bytes
const uint8_t kFrModulus[32] = {0x01, 0x00, 0x00, 0xf0, 0x93, 0xf5, 0xe1, 0x43, 0x91, 0x70, 0xb9,
                                0x79, 0x48, 0xe8, 0x33, 0x28, 0x5d, 0x58, 0x81, 0x81, 0xb6, 0x45,
                                0x50, 0xb8, 0x29, 0xa0, 0x31, 0xe1, 0x72, 0x4e, 0x64, 0x30};

struct Cursor {
  const uint8_t* p;
  const uint8_t* end;
  bool ok = true;
  bool need(uint64_t n) {
    if (!ok || uint64_t(end - p) < n) ok = false;
    return ok;
  }
  uint32_t u32() {
    if (!need(4)) return 0;
    uint32_t v;
    std::memcpy(&v, p, 4);
    p += 4;
    return v;
  }
  uint64_t u64() {
    if (!need(8)) return 0;
    uint64_t v;
    std::memcpy(&v, p, 8);
    p += 8;
    return v;
  }
  const uint8_t* bytes(uint64_t n) {
    if (!need(n)) return nullptr;
    const uint8_t* q = p;
    p += n;
    return q;
  }
};

bool canonical_fr(const uint8_t* v) {
  for (int i = 31; i >= 0; i--) {
    if (v[i] < kFrModulus[i]) return true;
    if (v[i] > kFrModulus[i]) return false;
  }
  return false;   // == r
}

struct Section {
  const uint8_t* p = nullptr;
  uint64_t size = 0;
};

// locate sections 1..max_type of a binfile; duplicate or truncated sections are errors
bool find_sections(const uint8_t* file, uint64_t len, const char magic[4], uint32_t min_version,
                   uint32_t max_version, Section* out, uint32_t max_type) {
  if (!file) return false;
  Cursor c{file, file + len};
  const uint8_t* m = c.bytes(4);
  if (!m || std::memcmp(m, magic, 4) != 0) return false;
  uint32_t version = c.u32(), n_sections = c.u32();
  if (!c.ok || version < min_version || version > max_version) return false;
  for (uint32_t i = 0; i < n_sections; i++) {
    uint32_t type = c.u32();
    uint64_t size = c.u64();
    const uint8_t* payload = c.bytes(size);
    if (!c.ok) return false;
    if (type >= 1 && type <= max_type) {
      if (out[type].p) return false;
      out[type] = {payload, size};
    }
  }
  return true;
}

struct R1csHeader {
  uint32_t n_wires, n_pub_out, n_pub_in, n_prv_in, n_constraints;
  uint64_t n_labels;
};

bool read_field_header(Cursor& c) {
  uint32_t n8 = c.u32();
  const uint8_t* prime = c.bytes(32);
  return c.ok && n8 == 32 && std::memcmp(prime, kFrModulus, 32) == 0;
}

bool r1cs_header(const Section& s, R1csHeader* h) {
  if (!s.p) return false;
  Cursor c{s.p, s.p + s.size};
  if (!read_field_header(c)) return false;
  h->n_wires = c.u32();
  h->n_pub_out = c.u32();
  h->n_pub_in = c.u32();
  h->n_prv_in = c.u32();
  h->n_labels = c.u64();
  h->n_constraints = c.u32();
  if (!c.ok || h->n_wires == 0) return false;
  return uint64_t(h->n_pub_out) + h->n_pub_in + h->n_prv_in + 1 <= h->n_wires;
}

// one walk over the constraint section; with null outputs it only validates and counts
bool r1cs_walk(const Section& s, const R1csHeader& h, uint64_t nnz[3], uint32_t* const row_ptr[3],
               uint32_t* const col[3], uint8_t* const coeff[3]) {
  if (!s.p && h.n_constraints) return false;
  Cursor c{s.p, s.p + s.size};
  uint64_t count[3] = {0, 0, 0};
  for (uint32_t j = 0; j < h.n_constraints; j++) {
    for (int k = 0; k < 3; k++) {
      if (row_ptr) row_ptr[k][j] = uint32_t(count[k]);
      uint32_t terms = c.u32();
      if (!c.need(uint64_t(terms) * 36)) return false;
      for (uint32_t t = 0; t < terms; t++) {
        uint32_t wire = c.u32();
        const uint8_t* v = c.bytes(32);
        if (wire >= h.n_wires || !canonical_fr(v)) return false;
        if (col) {
          col[k][count[k]] = wire;
          std::memcpy(coeff[k] + count[k] * 32, v, 32);
        }
        count[k]++;
      }
      if (count[k] >> 32) return false;   // CSR offsets are 32-bit
    }
  }
  if (!c.ok || c.p != c.end) return false;
  for (int k = 0; k < 3; k++) {
    if (row_ptr) row_ptr[k][h.n_constraints] = uint32_t(count[k]);
    if (nnz) nnz[k] = count[k];
  }
  return true;
}

}  // namespace

extern "C" int zkb_r1cs_inspect(const uint8_t* file, uint64_t len, zkb_r1cs_info* out) {
  if (!out) return ZKB_EINVAL;
  Section s[4];
  R1csHeader h;
  if (!find_sections(file, len, "r1cs", 1, 1, s, 3) || !r1cs_header(s[1], &h)) return ZKB_EINVAL;
  uint64_t nnz[3];
  if (!r1cs_walk(s[2], h, nnz, nullptr, nullptr, nullptr)) return ZKB_EINVAL;
  out->n_wires = h.n_wires;
  out->n_pub_out = h.n_pub_out;
  out->n_pub_in = h.n_pub_in;
  out->n_prv_in = h.n_prv_in;
  out->n_constraints = h.n_constraints;
  out->n_labels = h.n_labels;
  for (int k = 0; k < 3; k++) out->nnz[k] = nnz[k];
  return ZKB_OK;
}

extern "C" int zkb_r1cs_read(const uint8_t* file, uint64_t len, uint32_t* const row_ptr[3],
                             uint32_t* const col[3], uint8_t* const coeff[3]) {
  if (!row_ptr || !col || !coeff) return ZKB_EINVAL;
  for (int k = 0; k < 3; k++)
    if (!row_ptr[k]) return ZKB_EINVAL;
  Section s[4];
  R1csHeader h;
  if (!find_sections(file, len, "r1cs", 1, 1, s, 3) || !r1cs_header(s[1], &h)) return ZKB_EINVAL;
  // sizes were obtained from zkb_r1cs_inspect on the same bytes; col / coeff may be null only
  // for an empty matrix
  uint64_t nnz[3];
  if (!r1cs_walk(s[2], h, nnz, nullptr, nullptr, nullptr)) return ZKB_EINVAL;
  for (int k = 0; k < 3; k++)
    if (nnz[k] && (!col[k] || !coeff[k])) return ZKB_EINVAL;
  return r1cs_walk(s[2], h, nullptr, row_ptr, col, coeff) ? ZKB_OK : ZKB_EINVAL;
}

extern "C" int zkb_wtns_inspect(const uint8_t* file, uint64_t len, uint32_t* n_witness_out) {
  if (!n_witness_out) return ZKB_EINVAL;
  Section s[3];
  if (!find_sections(file, len, "wtns", 1, 2, s, 2) || !s[1].p) return ZKB_EINVAL;
  Cursor c{s[1].p, s[1].p + s[1].size};
  if (!read_field_header(c)) return ZKB_EINVAL;
  uint32_t n = c.u32();
  if (!c.ok || !s[2].p || s[2].size != uint64_t(n) * 32) return ZKB_EINVAL;
  *n_witness_out = n;
  return ZKB_OK;
}

extern "C" int zkb_wtns_read(const uint8_t* file, uint64_t len, uint8_t* witness_out) {
  uint32_t n = 0;
  if (!witness_out || zkb_wtns_inspect(file, len, &n) != ZKB_OK) return ZKB_EINVAL;
  Section s[3];
  find_sections(file, len, "wtns", 1, 2, s, 2);
  for (uint32_t i = 0; i < n; i++)
    if (!canonical_fr(s[2].p + uint64_t(i) * 32)) return ZKB_EINVAL;
  if (n == 0 || s[2].p[0] != 1) return ZKB_EINVAL;   // wire 0 is the constant 1
  for (int i = 1; i < 32; i++)
    if (s[2].p[i]) return ZKB_EINVAL;
  std::memcpy(witness_out, s[2].p, uint64_t(n) * 32);
  return ZKB_OK;
}

// ---------------------------------------------------------------- .zkey (snarkjs Groth16 proving key)
namespace {

using zkb::host::HFq;
using zkb::host::HFr;

// BN254 base field modulus q, little-endian bytes
const uint8_t kFqModulus[32] = {0x47, 0xfd, 0x7c, 0xd8, 0x16, 0x8c, 0x20, 0x3c, 0x8d, 0xca, 0x71,
                                0x68, 0x91, 0x6a, 0x81, 0x97, 0x5d, 0x58, 0x81, 0x81, 0xb6, 0x45,
                                0x50, 0xb8, 0x29, 0xa0, 0x31, 0xe1, 0x72, 0x4e, 0x64, 0x30};

struct ZkeyHeader {
  uint32_t n_vars, n_public, domain, log_n;
  const uint8_t* fixed;   // alpha_g1 beta_g1 beta_g2 gamma_g2 delta_g1 delta_g2 (64 64 128 128 64 128)
};

bool zkey_header(const Section* s, ZkeyHeader* h) {
  if (!s[1].p || !s[2].p) return false;
  Cursor c1{s[1].p, s[1].p + s[1].size};
  if (c1.u32() != 1 || !c1.ok) return false;                       // groth16
  Cursor c{s[2].p, s[2].p + s[2].size};
  uint32_t n8q = c.u32();
  const uint8_t* q = c.bytes(32);
  uint32_t n8r = c.u32();
  const uint8_t* r = c.bytes(32);
  if (!c.ok || n8q != 32 || n8r != 32 || std::memcmp(q, kFqModulus, 32) || std::memcmp(r, kFrModulus, 32))
    return false;
  h->n_vars = c.u32();
  h->n_public = c.u32();
  h->domain = c.u32();
  h->fixed = c.bytes(64 + 64 + 128 + 128 + 64 + 128);
  if (!c.ok || c.p != c.end) return false;
  if (h->domain < 2 || (h->domain & (h->domain - 1)) || h->n_vars == 0 || h->n_public >= h->n_vars) return false;
  h->log_n = 0;
  while ((1u << h->log_n) < h->domain) h->log_n++;
  // section sizes must agree with the header
  const uint64_t m = h->n_vars, np = uint64_t(h->n_public) + 1;
  return s[3].p && s[3].size == 64 * np && s[4].p && s[4].size >= 4 && s[5].p && s[5].size == 64 * m &&
         s[6].p && s[6].size == 64 * m && s[7].p && s[7].size == 128 * m && (s[8].p || m == np) &&
         s[8].size == 64 * (m - np) && s[9].p && s[9].size == 64ull * h->domain;
}

// count coordinates of 32 bytes each, Montgomery form -> canonical; false if one is not reduced
bool coords_from_mont(const uint8_t* in, uint64_t count, uint8_t* out) {
  for (uint64_t i = 0; i < count; i++) {
    HFq a;
    std::memcpy(a.l, in + 32 * i, 32);
    if (!HFq::lt_mod(a.l)) return false;
    a = a.from_mont();
    std::memcpy(out + 32 * i, a.l, 32);
  }
  return true;
}

// walks the coefficient section.  Pass 1 (col == nullptr): validates and counts per (matrix, row)
// into row_ptr[k][row + 1] when row_ptr is given, else only nnz / n_rows.  Pass 2: fills col / coeff
// using row_ptr[k][row] as the running cursor of the row.
bool zkey_coeffs(const Section& s, const ZkeyHeader& h, uint64_t nnz[2], uint32_t* n_rows,
                 uint32_t* const row_ptr[2], uint32_t* const col[2], uint8_t* const coeff[2]) {
  Cursor c{s.p, s.p + s.size};
  const uint32_t count = c.u32();
  if (!c.ok || s.size != 4 + 44ull * count) return false;
  uint64_t cnt[2] = {0, 0};
  uint32_t rows = 0;
  for (uint32_t i = 0; i < count; i++) {
    const uint32_t mat = c.u32(), row = c.u32(), sig = c.u32();
    const uint8_t* v = c.bytes(32);
    if (!c.ok || mat > 1 || row >= h.domain || sig >= h.n_vars) return false;
    if (row + 1 > rows) rows = row + 1;
    cnt[mat]++;
    if (col) {
      HFr a;
      std::memcpy(a.l, v, 32);
      if (!HFr::lt_mod(a.l)) return false;
      a = a.from_mont().from_mont();                                // value = coefficient * R^2
      const uint32_t pos = row_ptr[mat][row]++;
      col[mat][pos] = sig;
      std::memcpy(coeff[mat] + 32ull * pos, a.l, 32);
    } else if (row_ptr) {
      row_ptr[mat][row + 1]++;
    }
  }
  if (nnz) { nnz[0] = cnt[0]; nnz[1] = cnt[1]; }
  if (n_rows) *n_rows = rows;
  return true;
}

}  // namespace

extern "C" int zkb_zkey_inspect(const uint8_t* file, uint64_t len, zkb_zkey_info* out) {
  if (!out) return ZKB_EINVAL;
  Section s[11];
  ZkeyHeader h;
  if (!find_sections(file, len, "zkey", 1, 1, s, 10) || !zkey_header(s, &h)) return ZKB_EINVAL;
  uint64_t nnz[2];
  uint32_t rows;
  if (!zkey_coeffs(s[4], h, nnz, &rows, nullptr, nullptr, nullptr)) return ZKB_EINVAL;
  out->n_vars = h.n_vars;
  out->n_public = h.n_public;
  out->log_n = h.log_n;
  out->n_rows = rows;
  out->nnz[0] = nnz[0];
  out->nnz[1] = nnz[1];
  return ZKB_OK;
}

extern "C" int zkb_zkey_read(const uint8_t* file, uint64_t len, const zkb_zkey_data* d) {
  if (!d) return ZKB_EINVAL;
  Section s[11];
  ZkeyHeader h;
  if (!find_sections(file, len, "zkey", 1, 1, s, 10) || !zkey_header(s, &h)) return ZKB_EINVAL;
  uint64_t nnz[2];
  uint32_t rows;
  if (!zkey_coeffs(s[4], h, nnz, &rows, nullptr, nullptr, nullptr)) return ZKB_EINVAL;
  if (d->row_ptr[0] || d->row_ptr[1]) {
    // the matrices come together (one walk fills both)
    for (int k = 0; k < 2; k++)
      if (!d->row_ptr[k] || (nnz[k] && (!d->col[k] || !d->coeff[k]))) return ZKB_EINVAL;
    for (int k = 0; k < 2; k++) std::memset(d->row_ptr[k], 0, 4ull * (rows + 1));
    if (!zkey_coeffs(s[4], h, nullptr, nullptr, d->row_ptr, nullptr, nullptr)) return ZKB_EINVAL;
    for (int k = 0; k < 2; k++)                    // counts at [row + 1] -> row starts at [row]
      for (uint32_t r = 0; r < rows; r++) d->row_ptr[k][r + 1] += d->row_ptr[k][r];
    if (!zkey_coeffs(s[4], h, nullptr, nullptr, d->row_ptr, d->col, d->coeff)) return ZKB_EINVAL;
    for (int k = 0; k < 2; k++) {                  // the cursors ended at the next row's start: shift back
      for (uint32_t r = rows; r > 0; r--) d->row_ptr[k][r] = d->row_ptr[k][r - 1];
      d->row_ptr[k][0] = 0;
    }
  }
  const uint8_t* f = h.fixed;
  struct Item { uint8_t* out; const uint8_t* in; uint64_t coords; };
  const uint64_t m = h.n_vars, np = uint64_t(h.n_public) + 1;
  const Item items[] = {{d->alpha_g1, f, 2},       {d->beta_g1, f + 64, 2},   {d->beta_g2, f + 128, 4},
                        {d->gamma_g2, f + 256, 4}, {d->delta_g1, f + 384, 2}, {d->delta_g2, f + 448, 4},
                        {d->ic, s[3].p, 2 * np},   {d->a_query, s[5].p, 2 * m}, {d->b_g1_query, s[6].p, 2 * m},
                        {d->b_g2_query, s[7].p, 4 * m}, {d->c_query, s[8].p, 2 * (m - np)},
                        {d->h_query, s[9].p, 2ull * h.domain}};
  for (const Item& it : items)
    if (it.out && it.coords && !coords_from_mont(it.in, it.coords, it.out)) return ZKB_EINVAL;
  return ZKB_OK;
}
