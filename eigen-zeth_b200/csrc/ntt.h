// ntt.h — launch descriptors shared by ntt.cu (kernels) and the host planner in context.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace zkb {

void count_launches(uint64_t k);   // defined in msm_common.cu

struct NttPassArgs {
  const void* in;          // n x 32 B
  void* out;               // n x 32 B (== in for the in-place column passes)
  const void* stage_tw;    // 255 x 32 B dense per-stage twiddles (forward or inverse roots)
  const void* tw_lo;       // two-level inter-pass twiddles  w_n^i, i < 2^tw_lo_bits
  const void* tw_hi;       //                                w_n^{i * 2^tw_lo_bits}
  const void* tw_full;     // optional full table for this pass (M entries), else nullptr
  const void* pre_scale;   // optional n-entry table multiplied into the input (pass 0 only)
  const void* post_scale;  // optional n-entry table multiplied into the output (last pass only)
  uint64_t M;              // column mode: block size before this pass
  uint32_t S;              // column mode: M >> k
  uint32_t k;              // log2 of the sub-transform size of this pass (1..8)
  uint32_t tw_shift;       // log2(n / M)
  uint32_t tw_lo_bits;
  uint64_t tw_lo_mask;
  uint32_t row_bits;       // final mode: log2(number of rows) = log_n - k
  uint32_t first_digit_bits;  // final mode: width of K_0 (0 when P == 1)
  uint32_t n_row_digits;
  uint32_t row_digit_bits[4];
  uint32_t batch_log;      // each transform covers 2^batch_log elements; tiles beyond one transform's
  uint32_t tiles_per_batch;//   count belong to the next transform of a contiguous batch (final mode)
  // ---- fused exchange of the multi-GPU four-step NTT (dist.cu); both off for ordinary transforms
  // in_ilv: the FIRST pass reads a batch stored element-major: element i of transform b lives at
  //   i * batch + b (the layout the all-to-all delivers: [source row][local column]), so the unpack
  //   permutation is just this pass's load indexing.
  // out_win: the LAST pass stores element i of transform b to rank (i >> out_col_bits)'s receive
  //   window at row b, column i & (2^out_col_bits - 1): twiddle, pack and the all-to-all itself are
  //   this pass's store loop — out_wins[] are peer-mapped (NVLink) or local staging pointers.
  int in_ilv;
  uint32_t in_ilv_log;     // log2(batch)
  int out_win;
  uint32_t out_col_bits;
  uint32_t xform_log;      // log2 of the transform size (splits a logical index into (b, i))
  void* out_wins[16];
  // batch_cols (single-pass transforms only, P == 1): a tile's columns are CONSECUTIVE TRANSFORMS of a
  // contiguous batch (8 << (8 - k) of them per CTA) instead of one transform per CTA with the other
  // columns idle; logical index = transform * 2^k + element on both sides.
  int batch_cols;
  uint32_t batch_count;    // transforms in the batch (ragged last tile)
  uint32_t* flag;          // bit 0 raised when a canonical-form input (mont_in == 0) is >= r
  int mont_in;             // 0: input is canonical integers, convert while loading
  int mont_out;            // 0: convert the result back to canonical integers while storing
};

void ntt_launch_pass(const NttPassArgs& a, bool final, uint32_t tiles, cudaStream_t s);
void fr_powers_launch(void* out, const uint32_t base[8], const uint32_t c[8], uint64_t count,
                      uint64_t stride_pow, cudaStream_t s);
void fr_twiddle_full_launch(void* tab, const void* lo, const void* hi, uint32_t lo_bits,
                            uint32_t tw_shift, uint32_t s_bits, uint64_t M, cudaStream_t s);
void fr_convert_launch(void* data, uint64_t n, int to_mont, cudaStream_t s);
// index-mapped power tables for the distributed (four-step) NTT, pos = hi * cols + lo:
//   mode 0: out[pos] = c * base^{(row0 + hi) * lo}            (inter-step twiddle  w^{i2 * k1})
//   mode 1: out[pos] = c * base^{(row0 + hi) + lo * mul}      (coset powers at global index k1 + N1 k2)
void fr_table_launch(void* out, const uint32_t base[8], const uint32_t c[8], uint64_t rows, uint64_t cols,
                     uint64_t row0, uint64_t mul, int mode, cudaStream_t s);
void fr_gather_layout_launch(const void* full, void* out, uint32_t log_n, uint32_t d, uint32_t world,
                             uint32_t rank, cudaStream_t s);
// a[i] = a[i] * b[i] - c[i]
void fr_mul_sub_launch(void* a, const void* b, const void* c, uint64_t n, cudaStream_t s);

}  // namespace zkb
