"""ctypes binding of libzkb.so (include/zkb.h) — the host-side mirror of the C ABI.

There is NO CPU fallback: if the shared library is missing, or no sm_100 GPU is present,
every compute entry point raises.  Byte layouts are those of include/zkb.h (canonical
little-endian field elements; proof.json coordinate order of reference proof/proof.json).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libzkb.so")

ZKB_OK, ZKB_EINVAL, ZKB_ECUDA, ZKB_ENOMEM, ZKB_EBUSY, ZKB_ENOGPU, ZKB_EVERIFY = 0, -1, -2, -3, -4, -5, -6
VK_BYTES, PROOF_BYTES, INTERMEDIATE_BYTES = 576, 256, 384

EXPORTS = [
    "zkb_version", "zkb_ctx_create", "zkb_ctx_destroy", "zkb_last_error", "zkb_ctx_stream", "zkb_ctx_sync",
    "zkb_ctx_set", "zkb_fr_ntt", "zkb_fr_ntt_dev", "zkb_fr_ntt_batch_dev", "zkb_fr_table_dev",
    "zkb_fr_mul_sub_dev", "zkb_fr_convert_dev", "zkb_fr_root_of_unity", "zkb_bases_upload_g1", "zkb_bases_upload_g2",
    "zkb_bases_synthetic_g1", "zkb_bases_synthetic_g2", "zkb_bases_download", "zkb_bases_len", "zkb_bases_free",
    "zkb_bases_precompute", "zkb_bases_window",
    "zkb_msm_g1", "zkb_msm_g2", "zkb_msm_g1_dev", "zkb_msm_g2_dev", "zkb_g1_sum", "zkb_g2_sum",
    "zkb_pk_synthetic", "zkb_pk_synthetic_shard", "zkb_pk_load", "zkb_pk_free", "zkb_pk_query",
    "zkb_groth16_prove_partial", "zkb_groth16_assemble", "zkb_pk_synthetic_shard_ex", "zkb_groth16_witness_dev",
    "zkb_groth16_msm_dev", "zkb_fr_gather_layout_dev", "zkb_groth16_witness_layout_dev", "zkb_synthetic_witness", "zkb_groth16_prove", "zkb_groth16_verify", "zkb_groth16_verify_ex", "zkb_proof_check",
    "zkb_r1cs_inspect", "zkb_r1cs_read", "zkb_wtns_inspect", "zkb_wtns_read", "zkb_zkey_inspect", "zkb_zkey_read",
    "zkb_field_batch_op", "zkb_curve_batch_op", "zkb_bench_int_pipe", "zkb_g1_sum_dev", "zkb_g2_sum_dev",
    "zkb_ctx_launches", "zkb_msm_last_timings", "zkb_host_alloc", "zkb_host_free", "zkb_pk_n_vars", "zkb_pk_log_n",
    "zkb_comm_unique_id", "zkb_comm_init", "zkb_comm_init_local", "zkb_comm_rank", "zkb_comm_world",
    "zkb_comm_transport", "zkb_comm_barrier", "zkb_comm_destroy", "zkb_fr_ntt_dist_dev",
    "zkb_fr_index_power_table_dist", "zkb_msm_g1_dist_dev", "zkb_msm_g2_dist_dev", "zkb_msm_g1_dist",
    "zkb_msm_g2_dist", "zkb_pk_synthetic_dist", "zkb_groth16_prove_dist", "zkb_selftest_final_exp",
    "zkb_groth16_assemble_host", "zkb_selftest_assemble_tables",
]
COMM_ID_BYTES = 128


class ZkbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("zkb error %d: %s" % (code, msg))
        self.code = code


class PkDesc(C.Structure):
    _fields_ = [("log_n", C.c_uint32), ("n_public", C.c_uint32), ("n_vars", C.c_uint64), ("n_constraints", C.c_uint64),
                ("row_ptr", C.c_void_p * 3), ("col", C.c_void_p * 3), ("coeff", C.c_void_p * 3),
                ("a_query", C.c_void_p), ("b_g1_query", C.c_void_p), ("b_g2_query", C.c_void_p),
                ("l_query", C.c_void_p), ("h_query", C.c_void_p), ("alpha_g1", C.c_void_p), ("beta_g1", C.c_void_p),
                ("delta_g1", C.c_void_p), ("beta_g2", C.c_void_p), ("delta_g2", C.c_void_p),
                ("h_basis", C.c_uint32)]


class ProveTimings(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("ms_h2d", "ms_spmv", "ms_ntt", "ms_msm_a", "ms_msm_b1", "ms_msm_b2",
                                         "ms_msm_l", "ms_msm_h", "ms_total")]

    def as_dict(self):
        return {n: float(getattr(self, n)) for n, _ in self._fields_}


_lib = None


class R1csInfo(C.Structure):
    _fields_ = [("n_wires", C.c_uint32), ("n_pub_out", C.c_uint32), ("n_pub_in", C.c_uint32),
                ("n_prv_in", C.c_uint32), ("n_constraints", C.c_uint32), ("n_labels", C.c_uint64),
                ("nnz", C.c_uint64 * 3)]


class ZkeyInfo(C.Structure):
    _fields_ = [("n_vars", C.c_uint32), ("n_public", C.c_uint32), ("log_n", C.c_uint32), ("n_rows", C.c_uint32),
                ("nnz", C.c_uint64 * 2)]


class ZkeyData(C.Structure):
    _fields_ = [("row_ptr", C.c_void_p * 2), ("col", C.c_void_p * 2), ("coeff", C.c_void_p * 2)] + [
        (n, C.c_void_p) for n in ("alpha_g1", "beta_g1", "delta_g1", "beta_g2", "gamma_g2", "delta_g2", "ic",
                                  "a_query", "b_g1_query", "b_g2_query", "c_query", "h_query")]


def load_library() -> C.CDLL:
    """dlopen libzkb.so and declare prototypes; raises if the extension was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libzkb.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                          "there is no CPU fallback for the proving path")
    lib = C.CDLL(LIB_PATH)
    vp, u8p, u64, u32, i32 = C.c_void_p, C.c_char_p, C.c_uint64, C.c_uint32, C.c_int
    lib.zkb_version.restype = C.c_char_p
    lib.zkb_ctx_create.argtypes = [i32, C.POINTER(vp)]
    lib.zkb_ctx_destroy.argtypes = [vp]
    lib.zkb_ctx_destroy.restype = None
    lib.zkb_last_error.argtypes = [vp]
    lib.zkb_last_error.restype = C.c_char_p
    lib.zkb_ctx_stream.argtypes = [vp]
    lib.zkb_ctx_stream.restype = vp
    lib.zkb_ctx_sync.argtypes = [vp]
    lib.zkb_ctx_set.argtypes = [vp, C.c_char_p, C.c_int64]
    lib.zkb_host_alloc.argtypes = [u64, C.POINTER(vp)]
    lib.zkb_host_free.argtypes = [vp]
    lib.zkb_host_free.restype = None
    lib.zkb_fr_ntt.argtypes = [vp, vp, u32, i32, i32]
    lib.zkb_fr_ntt_dev.argtypes = [vp, vp, vp, u32, i32, i32, i32]
    lib.zkb_fr_ntt_batch_dev.argtypes = [vp, vp, vp, u32, u64, i32, i32, vp, vp]
    lib.zkb_fr_table_dev.argtypes = [vp, vp, u8p, u8p, u64, u64, u64, u64, i32]
    lib.zkb_fr_mul_sub_dev.argtypes = [vp, vp, vp, vp, u64]
    lib.zkb_fr_convert_dev.argtypes = [vp, vp, u64, i32]
    lib.zkb_fr_root_of_unity.argtypes = [u32, i32, vp]
    lib.zkb_fr_root_of_unity.restype = None
    for g in ("g1", "g2"):
        getattr(lib, "zkb_bases_upload_" + g).argtypes = [vp, u8p, u64, C.POINTER(vp)]
        getattr(lib, "zkb_bases_synthetic_" + g).argtypes = [vp, u64, u64, C.POINTER(vp)]
        getattr(lib, "zkb_msm_" + g).argtypes = [vp, vp, u64, vp, u64, vp]
        getattr(lib, "zkb_msm_" + g + "_dev").argtypes = [vp, vp, u64, vp, u64, vp]
        getattr(lib, "zkb_" + g + "_sum").argtypes = [vp, u8p, u32, vp]
    lib.zkb_bases_download.argtypes = [vp, vp, u64, u64, vp]
    lib.zkb_bases_len.argtypes = [vp]
    lib.zkb_bases_len.restype = u64
    lib.zkb_bases_precompute.argtypes = [vp, vp, i32]
    lib.zkb_bases_window.argtypes = [vp]
    lib.zkb_bases_free.argtypes = [vp]
    lib.zkb_bases_free.restype = None
    if hasattr(lib, "zkb_pk_synthetic"):
        lib.zkb_pk_synthetic.argtypes = [vp, u32, u64, C.POINTER(vp), vp]
        lib.zkb_pk_synthetic_shard.argtypes = [vp, u32, u64, u32, u32, C.POINTER(vp), vp]
        lib.zkb_pk_load.argtypes = [vp, C.POINTER(PkDesc), C.POINTER(vp)]
        lib.zkb_pk_synthetic_shard_ex.argtypes = [vp, u32, u64, u32, u32, u32, C.POINTER(vp), vp]
        lib.zkb_groth16_witness_dev.argtypes = [vp, vp, vp, C.POINTER(vp)]
        lib.zkb_groth16_msm_dev.argtypes = [vp, vp, vp, vp]
        lib.zkb_groth16_witness_layout_dev.argtypes = [vp, vp, vp, u32, u32, u32, C.POINTER(vp)]
        lib.zkb_fr_gather_layout_dev.argtypes = [vp, vp, vp, u32, u32, u32, u32]
        lib.zkb_groth16_prove_partial.argtypes = [vp, vp, vp, vp, C.POINTER(ProveTimings)]
        lib.zkb_groth16_assemble.argtypes = [vp, vp, u8p, u8p, u8p, vp]
        lib.zkb_pk_free.argtypes = [vp]
        lib.zkb_pk_free.restype = None
        lib.zkb_pk_query.argtypes = [vp, i32]
        lib.zkb_pk_query.restype = vp
        lib.zkb_pk_n_vars.argtypes = [vp]
        lib.zkb_pk_n_vars.restype = u64
        lib.zkb_pk_log_n.argtypes = [vp]
        lib.zkb_pk_log_n.restype = u32
        lib.zkb_synthetic_witness.argtypes = [u32, u8p, u8p, vp]
        lib.zkb_groth16_prove.argtypes = [vp, vp, vp, u8p, u8p, vp, vp, vp, C.POINTER(ProveTimings)]
        lib.zkb_groth16_verify.argtypes = [u8p, u8p, u8p]
        lib.zkb_groth16_verify_ex.argtypes = [u8p, u8p, u32, u8p, u8p]
        lib.zkb_proof_check.argtypes = [u8p, u8p, C.POINTER(u32)]
        lib.zkb_selftest_final_exp.argtypes = [u8p]
        lib.zkb_groth16_assemble_host.argtypes = [u8p, u8p, u8p, u8p, vp]
        lib.zkb_selftest_assemble_tables.argtypes = [u8p, u8p, u8p, u8p]
    lib.zkb_comm_unique_id.argtypes = [vp]
    lib.zkb_comm_init.argtypes = [vp, u8p, i32, i32]
    lib.zkb_comm_init_local.argtypes = [C.POINTER(vp), i32]
    for name in ("zkb_comm_rank", "zkb_comm_world", "zkb_comm_transport", "zkb_comm_barrier"):
        getattr(lib, name).argtypes = [vp]
    lib.zkb_comm_destroy.argtypes = [vp]
    lib.zkb_comm_destroy.restype = None
    lib.zkb_fr_ntt_dist_dev.argtypes = [vp, vp, u32, u32, i32, vp, vp]
    lib.zkb_fr_index_power_table_dist.argtypes = [vp, u32, u32, u8p, u8p, C.POINTER(vp)]
    for g in ("g1", "g2"):
        getattr(lib, "zkb_msm_%s_dist_dev" % g).argtypes = [vp, vp, u64, vp, u64, vp]
        getattr(lib, "zkb_msm_%s_dist" % g).argtypes = [vp, vp, u64, vp, u64, vp]
    lib.zkb_pk_synthetic_dist.argtypes = [vp, u32, u64, C.POINTER(vp), vp]
    lib.zkb_groth16_prove_dist.argtypes = [vp, vp, vp, u8p, u8p, vp, C.POINTER(ProveTimings)]
    lib.zkb_r1cs_inspect.argtypes = [u8p, u64, C.POINTER(R1csInfo)]
    lib.zkb_r1cs_read.argtypes = [u8p, u64, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    lib.zkb_wtns_inspect.argtypes = [u8p, u64, C.POINTER(u32)]
    lib.zkb_wtns_read.argtypes = [u8p, u64, vp]
    lib.zkb_zkey_inspect.argtypes = [u8p, u64, C.POINTER(ZkeyInfo)]
    lib.zkb_zkey_read.argtypes = [u8p, u64, C.POINTER(ZkeyData)]
    lib.zkb_field_batch_op.argtypes = [vp, i32, i32, u8p, u8p, vp, u64]
    lib.zkb_curve_batch_op.argtypes = [vp, i32, i32, u8p, u8p, vp, u64]
    lib.zkb_bench_int_pipe.argtypes = [vp, i32, C.POINTER(C.c_double)]
    lib.zkb_g1_sum_dev.argtypes = [vp, vp, u32, vp]
    lib.zkb_g2_sum_dev.argtypes = [vp, vp, u32, vp]
    lib.zkb_ctx_launches.argtypes = [vp]
    lib.zkb_ctx_launches.restype = u64
    lib.zkb_msm_last_timings.argtypes = [vp, C.POINTER(C.c_float)]
    _lib = lib
    return lib


def version() -> str:
    return load_library().zkb_version().decode()


class Bases:
    """device-resident MSM bases (affine, Montgomery form on the GPU)."""

    def __init__(self, ctx: "Context", handle, g2: bool):
        self.ctx, self.handle, self.g2 = ctx, handle, g2

    def __len__(self):
        return int(self.ctx.lib.zkb_bases_len(self.handle))

    def precompute(self, window_bits: int = 0) -> int:
        """expand to window multiples in HBM (zkb_bases_precompute); returns the window width."""
        self.ctx._check(self.ctx.lib.zkb_bases_precompute(self.ctx.h, self.handle, window_bits))
        return int(self.ctx.lib.zkb_bases_window(self.handle))

    @property
    def window(self) -> int:
        return int(self.ctx.lib.zkb_bases_window(self.handle))

    def download(self, first=0, count=None) -> bytes:
        count = len(self) - first if count is None else count
        buf = C.create_string_buffer(count * (128 if self.g2 else 64))
        self.ctx._check(self.ctx.lib.zkb_bases_download(self.ctx.h, self.handle, first, count, buf))
        return buf.raw

    def free(self):
        if self.handle:
            self.ctx.lib.zkb_bases_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class BorrowedBases(Bases):
    """view of bases owned by a proving key: never freed from here."""

    def free(self):
        self.handle = None


class ProvingKey:
    def __init__(self, ctx, handle, vk: bytes, log_n: int):
        self.ctx, self.handle, self.vk, self.log_n = ctx, handle, vk, log_n
        self.n_vars = int(ctx.lib.zkb_pk_n_vars(handle))

    def witness_ptr(self, witness):
        """bytes (length-checked against the key: the C ABI takes a bare pointer and would read
        32 * n_vars bytes whatever it is given) or a raw host / device address (the caller's promise)."""
        if isinstance(witness, (bytes, bytearray)):
            if len(witness) != 32 * self.n_vars:
                raise ValueError("witness holds %d bytes, the key has %d variables (%d bytes)"
                                 % (len(witness), self.n_vars, 32 * self.n_vars))
            buf = C.create_string_buffer(bytes(witness), len(witness))
            return C.cast(buf, C.c_void_p), buf
        return C.c_void_p(int(witness)), None

    def query(self, which: int) -> "Bases":
        """borrowed view of a query (0 a, 1 b_g1, 2 b_g2, 3 l, 4 h)."""
        h = self.ctx.lib.zkb_pk_query(self.handle, which)
        if not h:
            raise ValueError("no such query")
        return BorrowedBases(self.ctx, C.c_void_p(h), which == 2)

    def free(self):
        if self.handle:
            self.ctx.lib.zkb_pk_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """one GPU, one stream; not re-entrant (ZKB_EBUSY)."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.zkb_ctx_create(device, C.byref(h))
        if rc == ZKB_ENOGPU:
            raise ZkbError(rc, "no sm_100 GPU visible; the proving path has no CPU fallback")
        if rc:
            raise ZkbError(rc, "zkb_ctx_create failed")
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.zkb_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise ZkbError(rc, self.lib.zkb_last_error(self.h).decode(errors="replace"))

    @property
    def stream(self) -> int:
        return int(self.lib.zkb_ctx_stream(self.h) or 0)

    def sync(self):
        self._check(self.lib.zkb_ctx_sync(self.h))

    def set(self, key: str, value: int):
        self._check(self.lib.zkb_ctx_set(self.h, key.encode(), value))

    # ---- NTT
    def fr_ntt(self, data: bytes, inverse=False, coset=False) -> bytes:
        n = len(data) // 32
        if n * 32 != len(data) or n & (n - 1) or n < 2:
            raise ValueError("length must be 32 * 2^k, k >= 1")
        buf = C.create_string_buffer(bytes(data), len(data))
        self._check(self.lib.zkb_fr_ntt(self.h, buf, n.bit_length() - 1, int(inverse), int(coset)))
        return buf.raw

    def fr_ntt_dev(self, d_in: int, d_out: int, log_n: int, inverse=False, coset=False, montgomery=False):
        self._check(self.lib.zkb_fr_ntt_dev(self.h, d_in, d_out, log_n, int(inverse), int(coset), int(montgomery)))

    def fr_ntt_batch_dev(self, d_in, d_out, log_size, batch, inverse=False, skip_ninv=False, pre=0, post=0):
        self._check(self.lib.zkb_fr_ntt_batch_dev(self.h, d_in, d_out, log_size, batch, int(inverse), int(skip_ninv),
                                                  pre or None, post or None))

    def fr_table_dev(self, d_out, base: bytes, scale: bytes, rows, cols, row0, mul, mode):
        self._check(self.lib.zkb_fr_table_dev(self.h, d_out, base, scale, rows, cols, row0, mul, mode))

    def fr_mul_sub_dev(self, d_a, d_b, d_c, n):
        self._check(self.lib.zkb_fr_mul_sub_dev(self.h, d_a, d_b, d_c, n))

    def fr_convert_dev(self, d_data, n, to_montgomery: bool):
        self._check(self.lib.zkb_fr_convert_dev(self.h, d_data, n, int(to_montgomery)))

    # ---- bases / MSM
    def upload_bases(self, affine: bytes, g2=False) -> Bases:
        sz = 128 if g2 else 64
        n = len(affine) // sz
        h = C.c_void_p()
        fn = self.lib.zkb_bases_upload_g2 if g2 else self.lib.zkb_bases_upload_g1
        self._check(fn(self.h, bytes(affine), n, C.byref(h)))
        return Bases(self, h, g2)

    def synthetic_bases(self, n: int, seed: int, g2=False) -> Bases:
        h = C.c_void_p()
        fn = self.lib.zkb_bases_synthetic_g2 if g2 else self.lib.zkb_bases_synthetic_g1
        self._check(fn(self.h, n, seed, C.byref(h)))
        return Bases(self, h, g2)

    def msm(self, bases: Bases, scalars: bytes, first=0) -> bytes:
        n = len(scalars) // 32
        out = C.create_string_buffer(128 if bases.g2 else 64)
        fn = self.lib.zkb_msm_g2 if bases.g2 else self.lib.zkb_msm_g1
        sbuf = C.create_string_buffer(bytes(scalars), max(len(scalars), 1))
        self._check(fn(self.h, bases.handle, first, sbuf, n, out))
        return out.raw

    def msm_raw(self, bases: Bases, host_ptr: int, n: int, first=0) -> bytes:
        """scalars given as a raw host pointer (e.g. pinned torch tensor .data_ptr())."""
        out = C.create_string_buffer(128 if bases.g2 else 64)
        fn = self.lib.zkb_msm_g2 if bases.g2 else self.lib.zkb_msm_g1
        self._check(fn(self.h, bases.handle, first, host_ptr, n, out))
        return out.raw

    def msm_dev(self, bases: Bases, d_scalars: int, n: int, d_out: int, first=0):
        fn = self.lib.zkb_msm_g2_dev if bases.g2 else self.lib.zkb_msm_g1_dev
        self._check(fn(self.h, bases.handle, first, d_scalars, n, d_out))

    def point_sum(self, points: bytes, g2=False) -> bytes:
        sz = 128 if g2 else 64
        out = C.create_string_buffer(sz)
        fn = self.lib.zkb_g2_sum if g2 else self.lib.zkb_g1_sum
        self._check(fn(self.h, bytes(points), len(points) // sz, out))
        return out.raw

    def point_sum_dev(self, d_points: int, n: int, d_out: int, g2=False):
        fn = self.lib.zkb_g2_sum_dev if g2 else self.lib.zkb_g1_sum_dev
        self._check(fn(self.h, d_points, n, d_out))

    @property
    def launches(self) -> int:
        return int(self.lib.zkb_ctx_launches(self.h))

    def msm_last_timings(self):
        out = (C.c_float * 5)()
        self._check(self.lib.zkb_msm_last_timings(self.h, out))
        return dict(zip(("digits", "sort", "accumulate_l1", "accumulate_l2_finish", "reduce_tail"), [float(x) for x in out]))

    # ---- Groth16
    def pk_synthetic(self, log_n: int, seed: int) -> ProvingKey:
        h = C.c_void_p()
        vk = C.create_string_buffer(VK_BYTES)
        self._check(self.lib.zkb_pk_synthetic(self.h, log_n, seed, C.byref(h), vk))
        return ProvingKey(self, h, vk.raw, log_n)

    def pk_synthetic_shard(self, log_n: int, seed: int, rank: int, world: int, h_low_bits: int = 0) -> ProvingKey:
        h = C.c_void_p()
        vk = C.create_string_buffer(VK_BYTES)
        self._check(self.lib.zkb_pk_synthetic_shard_ex(self.h, log_n, seed, rank, world, h_low_bits, C.byref(h), vk))
        return ProvingKey(self, h, vk.raw, log_n)

    def pk_load(self, log_n: int, n_vars: int, n_public: int, matrices, queries: dict, fixed: dict, vk: bytes = b"",
                h_basis: int = 0) -> ProvingKey:
        """matrices: three (row_ptr list[int], col list[int], coeff bytes) for A, B, C;
        queries: a_query, b_g1_query, b_g2_query, l_query, h_query (canonical affine bytes);
        fixed: alpha_g1, beta_g1, delta_g1, beta_g2, delta_g2.
        h_basis = 1 (a snarkjs .zkey, see circom_io.read_zkey): matrices may hold A and B only, A
        includes the input rows, h_query holds 2^log_n points."""
        import array
        d = PkDesc()
        d.log_n, d.n_public, d.n_vars, d.h_basis = log_n, n_public, n_vars, h_basis
        if h_basis not in (0, 1) or len(matrices) not in ((3,) if h_basis == 0 else (2, 3)):
            raise ValueError("h_basis 0 takes A, B, C; h_basis 1 takes A, B and optionally C")
        d.n_constraints = len(matrices[0][0]) - 1
        # the C ABI takes bare pointers: every array length is checked here against the dimensions
        n = 1 << log_n
        want = {"a_query": 64 * n_vars, "b_g1_query": 64 * n_vars, "b_g2_query": 128 * n_vars,
                "l_query": 64 * max(n_vars - n_public - 1, 0), "h_query": 64 * (n if h_basis else n - 1)}
        for name, size in want.items():
            if len(queries[name]) != size:
                raise ValueError("%s holds %d bytes, expected %d" % (name, len(queries[name]), size))
        for name, size in (("alpha_g1", 64), ("beta_g1", 64), ("delta_g1", 64), ("beta_g2", 128), ("delta_g2", 128)):
            if len(fixed[name]) != size:
                raise ValueError("%s must be %d bytes" % (name, size))
        for k, (rp, col, coeff) in enumerate(matrices):
            if len(rp) != d.n_constraints + 1:
                raise ValueError("matrix %d: row_ptr needs n_constraints + 1 entries" % k)
            if len(col) != rp[-1] or len(coeff) != 32 * rp[-1]:
                raise ValueError("matrix %d: col / coeff lengths do not match row_ptr[-1] = %d" % (k, rp[-1]))
        keep = []
        for k, (rp, col, coeff) in enumerate(matrices):
            a_rp, a_col = array.array("I", rp), array.array("I", col)
            b_co = C.create_string_buffer(bytes(coeff), max(len(coeff), 1))
            keep += [a_rp, a_col, b_co]
            d.row_ptr[k] = a_rp.buffer_info()[0]
            d.col[k] = a_col.buffer_info()[0] if len(a_col) else None
            d.coeff[k] = C.cast(b_co, C.c_void_p)
        for name in ("a_query", "b_g1_query", "b_g2_query", "l_query", "h_query"):
            b = C.create_string_buffer(bytes(queries[name]), max(len(queries[name]), 1))
            keep.append(b)
            setattr(d, name, C.cast(b, C.c_void_p))
        for name in ("alpha_g1", "beta_g1", "delta_g1", "beta_g2", "delta_g2"):
            b = C.create_string_buffer(bytes(fixed[name]), len(fixed[name]))
            keep.append(b)
            setattr(d, name, C.cast(b, C.c_void_p))
        h = C.c_void_p()
        self._check(self.lib.zkb_pk_load(self.h, C.byref(d), C.byref(h)))
        return ProvingKey(self, h, vk, log_n)

    def groth16_witness_dev(self, pk: ProvingKey, witness):
        """H2D + R1CS products; returns the device pointers of the a, b, c evaluation vectors."""
        ptrs = (C.c_void_p * 3)()
        wptr, _keep = pk.witness_ptr(witness)
        self._check(self.lib.zkb_groth16_witness_dev(self.h, pk.handle, wptr, ptrs))
        return [int(p) for p in ptrs]

    def groth16_witness_layout_dev(self, pk: ProvingKey, witness, low_bits: int, world: int, rank: int, d_abc):
        """R1CS products for this rank's layout-d rows only, written into the three given buffers."""
        ptrs = (C.c_void_p * 3)(*[int(p) for p in d_abc])
        wptr, _keep = pk.witness_ptr(witness)
        self._check(self.lib.zkb_groth16_witness_layout_dev(self.h, pk.handle, wptr, low_bits, world, rank, ptrs))

    def groth16_msm_dev(self, pk: ProvingKey, d_h: int) -> bytes:
        out = C.create_string_buffer(INTERMEDIATE_BYTES)
        self._check(self.lib.zkb_groth16_msm_dev(self.h, pk.handle, d_h, out))
        return out.raw

    def fr_gather_layout_dev(self, d_full: int, d_out: int, log_n: int, low_bits: int, world: int, rank: int):
        self._check(self.lib.zkb_fr_gather_layout_dev(self.h, d_full, d_out, log_n, low_bits, world, rank))

    def groth16_prove_partial(self, pk: ProvingKey, witness):
        """-> (384-byte partial MSM sums over this key's point ranges, timings)"""
        out = C.create_string_buffer(INTERMEDIATE_BYTES)
        t = ProveTimings()
        wptr, _keep = pk.witness_ptr(witness)
        self._check(self.lib.zkb_groth16_prove_partial(self.h, pk.handle, wptr, out, C.byref(t)))
        return out.raw, t.as_dict()

    def groth16_assemble(self, pk: ProvingKey, sums: bytes, r: bytes, s: bytes) -> bytes:
        proof = C.create_string_buffer(PROOF_BYTES)
        self._check(self.lib.zkb_groth16_assemble(self.h, pk.handle, bytes(sums), r, s, proof))
        return proof.raw

    def groth16_prove(self, pk: ProvingKey, witness, r: bytes, s: bytes, want_intermediates=False, want_h=False):
        """witness: bytes or raw host pointer (int). returns (proof, intermediates|None, h|None, timings)."""
        proof = C.create_string_buffer(PROOF_BYTES)
        inter = C.create_string_buffer(INTERMEDIATE_BYTES) if want_intermediates else None
        hbuf = C.create_string_buffer(32 * ((1 << pk.log_n) - 1)) if want_h else None
        t = ProveTimings()
        wptr, _keep = pk.witness_ptr(witness)
        self._check(self.lib.zkb_groth16_prove(self.h, pk.handle, wptr, r, s, proof, inter, hbuf, C.byref(t)))
        return proof.raw, (inter.raw if inter else None), (hbuf.raw if hbuf else None), t.as_dict()

    # ---- several GPUs: collective calls over this context's communicator
    def comm_init(self, unique_id: bytes, rank: int, world: int):
        """join an NCCL communicator (one process or thread per rank); collective"""
        if len(unique_id) != COMM_ID_BYTES:
            raise ValueError("unique id must be %d bytes" % COMM_ID_BYTES)
        self._check(self.lib.zkb_comm_init(self.h, bytes(unique_id), rank, world))

    @property
    def comm_rank(self) -> int:
        return int(self.lib.zkb_comm_rank(self.h))

    @property
    def comm_world(self) -> int:
        return int(self.lib.zkb_comm_world(self.h))

    @property
    def comm_transport(self) -> str:
        return {0: "none", 1: "peer-memory windows (NVLink stores from the NTT epilogue)",
                2: "NCCL send/recv staging"}[int(self.lib.zkb_comm_transport(self.h))]

    def comm_barrier(self):
        self._check(self.lib.zkb_comm_barrier(self.h))

    def fr_ntt_dist_dev(self, d_data: int, log_n: int, low_bits: int, inverse=False, pre=0, post=0):
        self._check(self.lib.zkb_fr_ntt_dist_dev(self.h, d_data, log_n, low_bits, int(inverse), pre or None, post or None))

    def fr_index_power_table_dist(self, log_n: int, low_bits: int, base: int, scale: int = 1) -> int:
        out = C.c_void_p()
        from .proof_format import R
        self._check(self.lib.zkb_fr_index_power_table_dist(self.h, log_n, low_bits, (base % R).to_bytes(32, "little"),
                                                           (scale % R).to_bytes(32, "little"), C.byref(out)))
        return int(out.value)

    def msm_dist_dev(self, bases: Bases, d_scalars: int, n: int, d_out: int, first=0):
        fn = self.lib.zkb_msm_g2_dist_dev if bases.g2 else self.lib.zkb_msm_g1_dist_dev
        self._check(fn(self.h, bases.handle, first, d_scalars, n, d_out))

    def msm_dist_raw(self, bases: Bases, host_ptr: int, n: int, first=0) -> bytes:
        out = C.create_string_buffer(128 if bases.g2 else 64)
        fn = self.lib.zkb_msm_g2_dist if bases.g2 else self.lib.zkb_msm_g1_dist
        self._check(fn(self.h, bases.handle, first, host_ptr, n, out))
        return out.raw

    def pk_synthetic_dist(self, log_n: int, seed: int) -> ProvingKey:
        h = C.c_void_p()
        vk = C.create_string_buffer(VK_BYTES)
        self._check(self.lib.zkb_pk_synthetic_dist(self.h, log_n, seed, C.byref(h), vk))
        return ProvingKey(self, h, vk.raw, log_n)

    def groth16_prove_dist(self, pk: ProvingKey, witness, r: bytes, s: bytes):
        """collective: every rank calls it with the same witness; -> (proof, timings)"""
        proof = C.create_string_buffer(PROOF_BYTES)
        t = ProveTimings()
        wptr, _keep = pk.witness_ptr(witness)
        self._check(self.lib.zkb_groth16_prove_dist(self.h, pk.handle, wptr, r, s, proof, C.byref(t)))
        return proof.raw, t.as_dict()

    # ---- hooks
    def field_op(self, field: int, op: int, a: bytes, b: bytes | None = None) -> bytes:
        n = len(a) // 32
        out = C.create_string_buffer(max(len(a), 1))
        self._check(self.lib.zkb_field_batch_op(self.h, field, op, bytes(a), bytes(b) if b is not None else None, out, n))
        return out.raw[: len(a)]

    def curve_op(self, group: int, op: int, a: bytes, b: bytes | None = None) -> bytes:
        sz = 128 if group else 64
        n = len(a) // sz
        out = C.create_string_buffer(max(len(a), 1))
        self._check(self.lib.zkb_curve_batch_op(self.h, group, op, bytes(a), bytes(b) if b is not None else None, out, n))
        return out.raw[: len(a)]

    def bench_int_pipe(self, iters=2000):
        res = (C.c_double * 3)()
        self._check(self.lib.zkb_bench_int_pipe(self.h, iters, res))
        return {"imad_wide_per_s": res[0], "montmul_per_s": res[1], "sm_mhz": res[2]}


def comm_unique_id() -> bytes:
    """NCCL unique id for Context.comm_init (made by rank 0, handed to the other ranks)"""
    out = C.create_string_buffer(COMM_ID_BYTES)
    rc = load_library().zkb_comm_unique_id(out)
    if rc:
        raise ZkbError(rc, "zkb_comm_unique_id failed (is libnccl.so.2 loadable?)")
    return out.raw


class DeviceGroup:
    """All ranks of a communicator inside ONE process: one context and one host thread per GPU
    (zkb_comm_init_local: peer pointers + CUDA events, no NCCL).  This is how the prover service
    spreads one GenFinalProof over the GPUs of a box.  `devices` may repeat a device (tests on a
    single GPU)."""

    def __init__(self, devices):
        from concurrent.futures import ThreadPoolExecutor
        self.devices = list(devices)
        self.world = len(self.devices)
        self.ctxs = [Context(d) for d in self.devices]
        arr = (C.c_void_p * self.world)(*[c.h for c in self.ctxs])
        rc = load_library().zkb_comm_init_local(arr, self.world)
        if rc:
            raise ZkbError(rc, "zkb_comm_init_local failed (world must be a power of two <= 16; GPUs need peer access)")
        self.pool = [ThreadPoolExecutor(max_workers=1) for _ in self.devices]     # one thread per rank, for life
        self.pks, self.vk = None, None

    def run(self, fn):
        """fn(rank, ctx) on every rank concurrently (ctypes releases the GIL inside the library)"""
        futs = [self.pool[r].submit(fn, r, self.ctxs[r]) for r in range(self.world)]
        return [f.result() for f in futs]

    def pk_synthetic(self, log_n: int, seed: int):
        self.pks = self.run(lambda r, ctx: ctx.pk_synthetic_dist(log_n, seed))
        self.vk = self.pks[0].vk
        return self.pks

    def groth16_prove(self, witness, r: bytes, s: bytes):
        """one proof over all GPUs of the group -> (proof, timings of the slowest rank)"""
        out = self.run(lambda k, ctx: ctx.groth16_prove_dist(self.pks[k], witness, r, s))
        proofs = {p for p, _ in out}
        if len(proofs) != 1:
            raise RuntimeError("ranks disagree on the proof")
        return out[0][0], max((t for _, t in out), key=lambda t: t["ms_total"])

    def close(self):
        for p in self.pool:
            p.shutdown(wait=True)
        if self.pks:
            for pk in self.pks:
                pk.free()
        for c in self.ctxs:
            c.close()


def host_alloc(nbytes: int) -> int:
    """page-locked host memory (zkb_host_alloc) -> address"""
    p = C.c_void_p()
    rc = load_library().zkb_host_alloc(nbytes, C.byref(p))
    if rc:
        raise ZkbError(rc, "zkb_host_alloc(%d) failed" % nbytes)
    return int(p.value)


def host_free(ptr: int):
    load_library().zkb_host_free(C.c_void_p(ptr))


def host_read(ptr: int, nbytes: int) -> bytes:
    return C.string_at(ptr, nbytes)


def host_write(ptr: int, data: bytes):
    C.memmove(ptr, bytes(data), len(data))


def fr_root_of_unity(log_n: int, inverse=False) -> bytes:
    out = C.create_string_buffer(32)
    load_library().zkb_fr_root_of_unity(log_n, int(inverse), out)
    return out.raw


def synthetic_witness(log_n: int, w0: bytes, w1: bytes) -> bytes:
    lib = load_library()
    out = C.create_string_buffer(32 << log_n)
    rc = lib.zkb_synthetic_witness(log_n, w0, w1, out)
    if rc:
        raise ZkbError(rc, "zkb_synthetic_witness failed")
    return out.raw


def synthetic_witness_into(log_n: int, w0: bytes, w1: bytes, ptr: int):
    """as synthetic_witness, written straight into caller memory (e.g. a zkb_host_alloc buffer)"""
    rc = load_library().zkb_synthetic_witness(log_n, w0, w1, C.c_void_p(ptr))
    if rc:
        raise ZkbError(rc, "zkb_synthetic_witness failed")


def groth16_assemble_host(fixed: dict, sums: bytes, r: bytes, s: bytes) -> bytes:
    """five MSM sums + the key's fixed points (alpha_g1, beta_g1, delta_g1, beta_g2, delta_g2) + (r, s) -> proof"""
    blob = fixed["alpha_g1"] + fixed["beta_g1"] + fixed["delta_g1"] + fixed["beta_g2"] + fixed["delta_g2"]
    if len(blob) != 448 or len(sums) != INTERMEDIATE_BYTES:
        raise ValueError("fixed points must total 448 bytes, sums 384")
    out = C.create_string_buffer(PROOF_BYTES)
    rc = load_library().zkb_groth16_assemble_host(blob, bytes(sums), r, s, out)
    if rc:
        raise ZkbError(rc, "zkb_groth16_assemble_host: malformed input")
    return out.raw


def groth16_verify(vk: bytes, proof: bytes, public_input: bytes) -> bool:
    """CPU pairing check; False for an invalid proof, raises for a malformed one."""
    lib = load_library()
    rc = lib.zkb_groth16_verify(vk, proof, public_input)
    if rc == ZKB_OK:
        return True
    if rc == ZKB_EVERIFY:
        return False
    raise ZkbError(rc, "malformed verifying key / proof / input")


def groth16_verify_ex(vk_fixed: bytes, ic: bytes, proof: bytes, public_inputs: bytes) -> bool:
    """the pairing check for any number of public inputs: vk_fixed = alpha_g1 | beta_g2 | gamma_g2 |
    delta_g2 (448 bytes), ic = n_public + 1 G1 points, public_inputs = n_public canonical Fr"""
    if len(vk_fixed) != 448 or len(ic) % 64 or len(ic) < 64 or len(public_inputs) != 32 * (len(ic) // 64 - 1) or len(proof) != PROOF_BYTES:
        raise ValueError("vk_fixed 448 bytes, ic 64 * (n_public + 1), inputs 32 * n_public, proof 256")
    rc = load_library().zkb_groth16_verify_ex(vk_fixed, ic, len(ic) // 64 - 1, proof, public_inputs or None)
    if rc == ZKB_OK:
        return True
    if rc == ZKB_EVERIFY:
        return False
    raise ZkbError(rc, "malformed verifying key / proof / input")


PROOF_CHECK_REASONS = {
    0x001: "pi_a coordinate >= q", 0x002: "pi_a not on the curve", 0x004: "pi_b coordinate >= q",
    0x008: "pi_b not on the twist", 0x010: "pi_b outside the order-r subgroup", 0x020: "pi_c coordinate >= q",
    0x040: "pi_c not on the curve", 0x080: "public input >= r", 0x100: "proof element is the point at infinity",
}


def proof_check(proof: bytes, public_input: bytes) -> list:
    """strict well-formedness check (zkb_proof_check): [] when fine, else the list of reasons."""
    if len(proof) != 256 or len(public_input) != 32:
        raise ValueError("proof is 256 bytes, public input 32")
    bits = C.c_uint32(0)
    load_library().zkb_proof_check(proof, public_input, C.byref(bits))
    return [msg for bit, msg in PROOF_CHECK_REASONS.items() if bits.value & bit]
