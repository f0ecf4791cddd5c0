//! zkb-sys — raw FFI bindings to libzkb.so (include/zkb.h), the B200-native BN254 Groth16
//! proving backend behind eigen-zeth's ProverService.
//!
//! NOT BUILT IN THIS REPOSITORY'S IMAGE (no Rust toolchain): this crate is the
//! binding a maintainer of eigen-zeth adds next to the reference's own build.rs.  It is generated
//! from include/zkb.h (tools/gen_rust_sys.py) and tests/test_abi_host.py checks that every
//! symbol the header declares appears here and is exported by the shared library.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)] pub struct zkb_ctx { _private: [u8; 0] }
#[repr(C)] pub struct zkb_bases { _private: [u8; 0] }
#[repr(C)] pub struct zkb_pk { _private: [u8; 0] }

pub const ZKB_OK: c_int = 0;
pub const ZKB_EINVAL: c_int = -1;
pub const ZKB_ECUDA: c_int = -2;
pub const ZKB_ENOMEM: c_int = -3;
pub const ZKB_EBUSY: c_int = -4;
pub const ZKB_ENOGPU: c_int = -5;
pub const ZKB_EVERIFY: c_int = -6;
pub const ZKB_ECOMM: c_int = -7;
pub const ZKB_VK_BYTES: usize = 576;
pub const ZKB_PROOF_BYTES: usize = 256;
pub const ZKB_INTERMEDIATE_BYTES: usize = 384;
pub const ZKB_COMM_ID_BYTES: usize = 128;
// reason bits of zkb_proof_check
pub const ZKB_BAD_A_RANGE: u32 = 0x001;
pub const ZKB_BAD_A_CURVE: u32 = 0x002;
pub const ZKB_BAD_B_RANGE: u32 = 0x004;
pub const ZKB_BAD_B_CURVE: u32 = 0x008;
pub const ZKB_BAD_B_SUBGROUP: u32 = 0x010;
pub const ZKB_BAD_C_RANGE: u32 = 0x020;
pub const ZKB_BAD_C_CURVE: u32 = 0x040;
pub const ZKB_BAD_INPUT_RANGE: u32 = 0x080;
pub const ZKB_BAD_INFINITY: u32 = 0x100;

#[repr(C)]
#[derive(Debug, Default, Clone, Copy)]
pub struct zkb_prove_timings {
    pub ms_h2d: f32, pub ms_spmv: f32, pub ms_ntt: f32, pub ms_msm_a: f32, pub ms_msm_b1: f32,
    pub ms_msm_b2: f32, pub ms_msm_l: f32, pub ms_msm_h: f32, pub ms_total: f32,
}

/// external proving key + R1CS (see include/zkb.h `zkb_pk_desc`)
#[repr(C)]
pub struct zkb_pk_desc {
    pub log_n: u32,
    pub n_public: u32,
    pub n_vars: u64,
    pub n_constraints: u64,
    pub row_ptr: [*const u32; 3],
    pub col: [*const u32; 3],
    pub coeff: [*const u8; 3],
    pub a_query: *const u8, pub b_g1_query: *const u8, pub b_g2_query: *const u8,
    pub l_query: *const u8, pub h_query: *const u8,
    pub alpha_g1: *const u8, pub beta_g1: *const u8, pub delta_g1: *const u8,
    pub beta_g2: *const u8, pub delta_g2: *const u8,
    /// 0 = ark-groth16 h_query, 1 = snarkjs .zkey H section (see include/zkb.h)
    pub h_basis: u32,
}

/// sizes of an iden3 .r1cs file (see include/zkb.h `zkb_r1cs_info`)
#[repr(C)]
#[derive(Debug, Default, Clone, Copy)]
pub struct zkb_r1cs_info {
    pub n_wires: u32, pub n_pub_out: u32, pub n_pub_in: u32, pub n_prv_in: u32, pub n_constraints: u32,
    pub n_labels: u64,
    pub nnz: [u64; 3],
}

/// sizes of a snarkjs Groth16 .zkey file (see include/zkb.h `zkb_zkey_info`)
#[repr(C)]
#[derive(Debug, Default, Clone, Copy)]
pub struct zkb_zkey_info {
    pub n_vars: u32, pub n_public: u32, pub log_n: u32, pub n_rows: u32,
    pub nnz: [u64; 2],
}

/// caller-allocated outputs of `zkb_zkey_read` (null pointers are skipped)
#[repr(C)]
pub struct zkb_zkey_data {
    pub row_ptr: [*mut u32; 2],
    pub col: [*mut u32; 2],
    pub coeff: [*mut u8; 2],
    pub alpha_g1: *mut u8, pub beta_g1: *mut u8, pub delta_g1: *mut u8,
    pub beta_g2: *mut u8, pub gamma_g2: *mut u8, pub delta_g2: *mut u8, pub ic: *mut u8,
    pub a_query: *mut u8, pub b_g1_query: *mut u8, pub b_g2_query: *mut u8,
    pub c_query: *mut u8, pub h_query: *mut u8,
}

#[link(name = "zkb")]
extern "C" {
    pub fn zkb_ctx_create(device: c_int, out_: *mut *mut zkb_ctx) -> c_int;
    pub fn zkb_ctx_destroy(ctx: *mut zkb_ctx);
    pub fn zkb_last_error(ctx: *const zkb_ctx) -> *const c_char;
    pub fn zkb_ctx_stream(ctx: *mut zkb_ctx) -> *mut c_void;
    pub fn zkb_ctx_sync(ctx: *mut zkb_ctx) -> c_int;
    pub fn zkb_ctx_set(ctx: *mut zkb_ctx, key: *const c_char, value: i64) -> c_int;
    pub fn zkb_version() -> *const c_char;
    pub fn zkb_host_alloc(bytes: u64, out_: *mut *mut c_void) -> c_int;
    pub fn zkb_host_free(p: *mut c_void);
    pub fn zkb_fr_ntt(ctx: *mut zkb_ctx, data: *mut u8, log_n: u32, inverse: c_int, coset: c_int) -> c_int;
    pub fn zkb_fr_ntt_dev(ctx: *mut zkb_ctx, d_in: *const c_void, d_out: *mut c_void, log_n: u32, inverse: c_int, coset: c_int, montgomery: c_int) -> c_int;
    pub fn zkb_fr_ntt_batch_dev(ctx: *mut zkb_ctx, d_in: *const c_void, d_out: *mut c_void, log_size: u32, batch: u64, inverse: c_int, skip_ninv: c_int, d_pre_table: *const c_void, d_post_table: *const c_void) -> c_int;
    pub fn zkb_fr_gather_layout_dev(ctx: *mut zkb_ctx, d_full: *const c_void, d_out: *mut c_void, log_n: u32, low_bits: u32, world: u32, rank: u32) -> c_int;
    pub fn zkb_fr_table_dev(ctx: *mut zkb_ctx, d_out: *mut c_void, base: *const u8, scale: *const u8, rows: u64, cols: u64, row0: u64, mul: u64, mode: c_int) -> c_int;
    pub fn zkb_fr_mul_sub_dev(ctx: *mut zkb_ctx, d_a: *mut c_void, d_b: *const c_void, d_c: *const c_void, n: u64) -> c_int;
    pub fn zkb_fr_convert_dev(ctx: *mut zkb_ctx, d_data: *mut c_void, n: u64, to_montgomery: c_int) -> c_int;
    pub fn zkb_fr_root_of_unity(log_n: u32, inverse: c_int, out_: *mut u8);
    pub fn zkb_bases_upload_g1(ctx: *mut zkb_ctx, affine: *const u8, n: u64, out_: *mut *mut zkb_bases) -> c_int;
    pub fn zkb_bases_upload_g2(ctx: *mut zkb_ctx, affine: *const u8, n: u64, out_: *mut *mut zkb_bases) -> c_int;
    pub fn zkb_bases_synthetic_g1(ctx: *mut zkb_ctx, n: u64, seed: u64, out_: *mut *mut zkb_bases) -> c_int;
    pub fn zkb_bases_synthetic_g2(ctx: *mut zkb_ctx, n: u64, seed: u64, out_: *mut *mut zkb_bases) -> c_int;
    pub fn zkb_bases_download(ctx: *mut zkb_ctx, b: *const zkb_bases, first: u64, count: u64, out_: *mut u8) -> c_int;
    pub fn zkb_bases_len(b: *const zkb_bases) -> u64;
    pub fn zkb_bases_precompute(ctx: *mut zkb_ctx, b: *mut zkb_bases, window_bits: c_int) -> c_int;
    pub fn zkb_bases_window(b: *const zkb_bases) -> c_int;
    pub fn zkb_bases_free(b: *mut zkb_bases);
    pub fn zkb_msm_g1(ctx: *mut zkb_ctx, b: *const zkb_bases, first: u64, scalars: *const u8, n: u64, out_: *mut u8) -> c_int;
    pub fn zkb_msm_g2(ctx: *mut zkb_ctx, b: *const zkb_bases, first: u64, scalars: *const u8, n: u64, out_: *mut u8) -> c_int;
    pub fn zkb_msm_g1_dev(ctx: *mut zkb_ctx, b: *const zkb_bases, first: u64, d_scalars: *const c_void, n: u64, d_out: *mut c_void) -> c_int;
    pub fn zkb_msm_g2_dev(ctx: *mut zkb_ctx, b: *const zkb_bases, first: u64, d_scalars: *const c_void, n: u64, d_out: *mut c_void) -> c_int;
    pub fn zkb_g1_sum(ctx: *mut zkb_ctx, points: *const u8, n: u32, out_: *mut u8) -> c_int;
    pub fn zkb_g2_sum(ctx: *mut zkb_ctx, points: *const u8, n: u32, out_: *mut u8) -> c_int;
    pub fn zkb_g1_sum_dev(ctx: *mut zkb_ctx, d_points: *const c_void, n: u32, d_out: *mut c_void) -> c_int;
    pub fn zkb_g2_sum_dev(ctx: *mut zkb_ctx, d_points: *const c_void, n: u32, d_out: *mut c_void) -> c_int;
    pub fn zkb_comm_unique_id(id_out: *mut u8) -> c_int;
    pub fn zkb_comm_init(ctx: *mut zkb_ctx, id: *const u8, rank: c_int, world: c_int) -> c_int;
    pub fn zkb_comm_init_local(ctxs: *mut *mut zkb_ctx, world: c_int) -> c_int;
    pub fn zkb_comm_rank(ctx: *const zkb_ctx) -> c_int;
    pub fn zkb_comm_world(ctx: *const zkb_ctx) -> c_int;
    pub fn zkb_comm_transport(ctx: *const zkb_ctx) -> c_int;
    pub fn zkb_comm_barrier(ctx: *mut zkb_ctx) -> c_int;
    pub fn zkb_comm_destroy(ctx: *mut zkb_ctx);
    pub fn zkb_fr_ntt_dist_dev(ctx: *mut zkb_ctx, d_data: *mut c_void, log_n: u32, low_bits: u32, inverse: c_int, d_pre_table: *const c_void, d_post_table: *const c_void) -> c_int;
    pub fn zkb_fr_index_power_table_dist(ctx: *mut zkb_ctx, log_n: u32, low_bits: u32, base: *const u8, scale: *const u8, d_table_out: *const *mut c_void) -> c_int;
    pub fn zkb_msm_g1_dist_dev(ctx: *mut zkb_ctx, b: *const zkb_bases, first: u64, d_scalars: *const c_void, n: u64, d_out: *mut c_void) -> c_int;
    pub fn zkb_msm_g2_dist_dev(ctx: *mut zkb_ctx, b: *const zkb_bases, first: u64, d_scalars: *const c_void, n: u64, d_out: *mut c_void) -> c_int;
    pub fn zkb_msm_g1_dist(ctx: *mut zkb_ctx, b: *const zkb_bases, first: u64, scalars: *const u8, n: u64, out_: *mut u8) -> c_int;
    pub fn zkb_msm_g2_dist(ctx: *mut zkb_ctx, b: *const zkb_bases, first: u64, scalars: *const u8, n: u64, out_: *mut u8) -> c_int;
    pub fn zkb_pk_synthetic(ctx: *mut zkb_ctx, log_n: u32, seed: u64, out_: *mut *mut zkb_pk, vk_out: *mut u8) -> c_int;
    pub fn zkb_pk_synthetic_shard(ctx: *mut zkb_ctx, log_n: u32, seed: u64, rank: u32, world: u32, out_: *mut *mut zkb_pk, vk_out: *mut u8) -> c_int;
    pub fn zkb_pk_load(ctx: *mut zkb_ctx, desc: *const zkb_pk_desc, out_: *mut *mut zkb_pk) -> c_int;
    pub fn zkb_pk_synthetic_shard_ex(ctx: *mut zkb_ctx, log_n: u32, seed: u64, rank: u32, world: u32, h_low_bits: u32, out_: *mut *mut zkb_pk, vk_out: *mut u8) -> c_int;
    pub fn zkb_pk_free(pk: *mut zkb_pk);
    pub fn zkb_pk_query(pk: *const zkb_pk, which: c_int) -> *const zkb_bases;
    pub fn zkb_pk_n_vars(pk: *const zkb_pk) -> u64;
    pub fn zkb_pk_log_n(pk: *const zkb_pk) -> u32;
    pub fn zkb_synthetic_witness(log_n: u32, w0: *const u8, w1: *const u8, z_out: *mut u8) -> c_int;
    pub fn zkb_groth16_prove(ctx: *mut zkb_ctx, pk: *const zkb_pk, witness: *const u8, r: *const u8, s: *const u8, proof_out: *mut u8, intermediates_opt: *mut u8, h_out_opt: *mut u8, timings_opt: *mut zkb_prove_timings) -> c_int;
    pub fn zkb_groth16_prove_partial(ctx: *mut zkb_ctx, pk: *const zkb_pk, witness: *const u8, partial_out: *mut u8, timings_opt: *mut zkb_prove_timings) -> c_int;
    pub fn zkb_groth16_witness_dev(ctx: *mut zkb_ctx, pk: *const zkb_pk, witness: *const u8, d_abc_out: *mut *mut c_void) -> c_int;
    pub fn zkb_groth16_witness_layout_dev(ctx: *mut zkb_ctx, pk: *const zkb_pk, witness: *const u8, low_bits: u32, world: u32, rank: u32, d_abc_local: *mut *mut c_void) -> c_int;
    pub fn zkb_groth16_msm_dev(ctx: *mut zkb_ctx, pk: *const zkb_pk, d_h: *const c_void, partial_out: *mut u8) -> c_int;
    pub fn zkb_groth16_assemble(ctx: *mut zkb_ctx, pk: *const zkb_pk, sums: *const u8, r: *const u8, s: *const u8, proof_out: *mut u8) -> c_int;
    pub fn zkb_pk_synthetic_dist(ctx: *mut zkb_ctx, log_n: u32, seed: u64, out_: *mut *mut zkb_pk, vk_out: *mut u8) -> c_int;
    pub fn zkb_groth16_prove_dist(ctx: *mut zkb_ctx, pk: *const zkb_pk, witness: *const u8, r: *const u8, s: *const u8, proof_out: *mut u8, timings_opt: *mut zkb_prove_timings) -> c_int;
    pub fn zkb_groth16_assemble_host(fixed: *const u8, sums: *const u8, r: *const u8, s: *const u8, proof_out: *mut u8) -> c_int;
    pub fn zkb_groth16_verify(vk: *const u8, proof: *const u8, public_input: *const u8) -> c_int;
    pub fn zkb_groth16_verify_ex(vk_fixed: *const u8, ic: *const u8, n_public: u32, proof: *const u8, public_inputs: *const u8) -> c_int;
    pub fn zkb_proof_check(proof: *const u8, public_input: *const u8, reasons_opt: *mut u32) -> c_int;
    pub fn zkb_selftest_final_exp(proof: *const u8) -> c_int;
    pub fn zkb_selftest_assemble_tables(fixed: *const u8, sums: *const u8, r: *const u8, s: *const u8) -> c_int;
    pub fn zkb_r1cs_inspect(file: *const u8, len: u64, out_: *mut zkb_r1cs_info) -> c_int;
    pub fn zkb_r1cs_read(file: *const u8, len: u64, row_ptr: *mut *mut u32, col: *mut *mut u32, coeff: *mut *mut u8) -> c_int;
    pub fn zkb_wtns_inspect(file: *const u8, len: u64, n_witness_out: *mut u32) -> c_int;
    pub fn zkb_wtns_read(file: *const u8, len: u64, witness_out: *mut u8) -> c_int;
    pub fn zkb_zkey_inspect(file: *const u8, len: u64, out_: *mut zkb_zkey_info) -> c_int;
    pub fn zkb_zkey_read(file: *const u8, len: u64, out_: *const zkb_zkey_data) -> c_int;
    pub fn zkb_field_batch_op(ctx: *mut zkb_ctx, field: c_int, op: c_int, a: *const u8, b: *const u8, out_: *mut u8, n: u64) -> c_int;
    pub fn zkb_curve_batch_op(ctx: *mut zkb_ctx, group: c_int, op: c_int, a: *const u8, b: *const u8, out_: *mut u8, n: u64) -> c_int;
    pub fn zkb_bench_int_pipe(ctx: *mut zkb_ctx, iters: c_int, results: *mut f64) -> c_int;
    pub fn zkb_ctx_launches(ctx: *const zkb_ctx) -> u64;
    pub fn zkb_msm_last_timings(ctx: *mut zkb_ctx, ms_out: *mut f32) -> c_int;
}
