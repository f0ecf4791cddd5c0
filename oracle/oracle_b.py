"""ctypes wrapper of Oracle B (oracle/oracle_b.cpp).  TEST INFRASTRUCTURE ONLY — see the header of
oracle_b.cpp.  Imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs; never by the
product package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "build", "liboracle_b.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "oracle_b.cpp")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True, capture_output=True)
    return _LIB


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        l = C.CDLL(_LIB)
        vp, u64 = C.c_void_p, C.c_uint64
        l.ob_threads.restype = C.c_int
        l.ob_fr_ntt.argtypes = [vp, C.c_uint, C.c_int, C.c_int]
        l.ob_fr_ntt.restype = None
        l.ob_fr_ntt_timed.argtypes = [vp, C.c_uint, C.c_int, C.c_int, C.c_int]
        l.ob_fr_ntt_timed.restype = C.c_double
        for g in ("g1", "g2"):
            f = getattr(l, "ob_msm_" + g)
            f.argtypes = [vp, vp, u64, vp]
            f.restype = C.c_double
            f = getattr(l, "ob_synthetic_bases_" + g)
            f.argtypes = [u64, u64, vp]
            f.restype = None
        l.ob_set_threads.argtypes = [C.c_int]
        l.ob_set_threads.restype = None
        l.ob_fr_dot.argtypes = [vp, vp, u64, vp]
        l.ob_fr_dot.restype = None
        l.ob_fr_sum.argtypes = [vp, u64, vp]
        l.ob_fr_sum.restype = None
        l.ob_g1_mul_gen.argtypes = [C.c_char_p, vp]
        l.ob_g1_mul_gen.restype = None
        l.ob_g2_mul_gen.argtypes = [C.c_char_p, vp]
        l.ob_g2_mul_gen.restype = None
        l.ob_random_scalars.argtypes = [u64, u64, vp]
        l.ob_random_scalars.restype = None
        l.ob_setup_synthetic.argtypes = [C.c_uint, u64]
        l.ob_setup_synthetic.restype = vp
        l.ob_pk_free.argtypes = [vp]
        l.ob_pk_free.restype = None
        l.ob_pk_vk.argtypes = [vp, vp]
        l.ob_pk_vk.restype = None
        l.ob_pk_fixed.argtypes = [vp, vp]
        l.ob_pk_fixed.restype = None
        l.ob_pk_query.argtypes = [vp, C.c_int, vp]
        l.ob_pk_query.restype = u64
        l.ob_synthetic_witness.argtypes = [C.c_uint, C.c_char_p, C.c_char_p, vp]
        l.ob_synthetic_witness.restype = None
        l.ob_groth16_prove.argtypes = [vp, vp, C.c_char_p, C.c_char_p, vp, vp, vp]
        l.ob_groth16_prove.restype = C.c_double
        _lib = l
    return _lib


def threads() -> int:
    return lib().ob_threads()


def set_threads(n: int = 0) -> int:
    """use n OpenMP threads (0 = every core of the host), whatever OMP_NUM_THREADS says: torchrun
    exports OMP_NUM_THREADS=1 to its workers, which would silently make the CPU baseline scalar."""
    lib().ob_set_threads(n if n > 0 else (os.cpu_count() or 1))
    return threads()


def _ptr(data):
    """bytes / bytearray / ctypes buffer / raw address -> something ctypes passes as void* without copying"""
    if isinstance(data, int):
        return C.c_void_p(data)
    if isinstance(data, bytes):
        return C.cast(C.c_char_p(data), C.c_void_p)
    if isinstance(data, bytearray):
        return C.cast((C.c_char * len(data)).from_buffer(data), C.c_void_p)
    return C.cast(data, C.c_void_p)


def fr_dot(a, b, n: int) -> bytes:
    """sum a_i * b_i mod r over two arrays of n canonical scalars (bytes or raw host addresses)."""
    out = C.create_string_buffer(32)
    lib().ob_fr_dot(_ptr(a), _ptr(b), n, out)
    return out.raw


def fr_sum(a, n: int) -> bytes:
    out = C.create_string_buffer(32)
    lib().ob_fr_sum(_ptr(a), n, out)
    return out.raw


def g1_mul_gen(k: bytes) -> bytes:
    out = C.create_string_buffer(64)
    lib().ob_g1_mul_gen(bytes(k), out)
    return out.raw


def g2_mul_gen(k: bytes) -> bytes:
    out = C.create_string_buffer(128)
    lib().ob_g2_mul_gen(bytes(k), out)
    return out.raw


def msm_closed_form(scalars, n: int, bases_seed: int, g2=False) -> bytes:
    """sum_i s_i * (k_i * G) for the synthetic bases of `bases_seed` (k_i = xoshiro256**(seed)),
    computed as (sum s_i k_i mod r) * G: the size-independent check of a large MSM."""
    ks = random_scalars(n, bases_seed)
    d = fr_dot(scalars, ks, n)
    return g2_mul_gen(d) if g2 else g1_mul_gen(d)


def _buf(data: bytes):
    return C.create_string_buffer(bytes(data), max(len(data), 1))


def fr_ntt(data: bytes, inverse=False, coset=False) -> bytes:
    n = len(data) // 32
    b = _buf(data)
    lib().ob_fr_ntt(b, n.bit_length() - 1, int(inverse), int(coset))
    return b.raw[: len(data)]


def fr_ntt_timed(data: bytes, inverse=False, coset=False, reps=1):
    n = len(data) // 32
    b = _buf(data)
    t = lib().ob_fr_ntt_timed(b, n.bit_length() - 1, int(inverse), int(coset), reps)
    return b.raw[: len(data)], t


def msm_g1(bases: bytes, scalars: bytes):
    n = len(scalars) // 32
    out = C.create_string_buffer(64)
    t = lib().ob_msm_g1(_buf(bases), _buf(scalars), n, out)
    return out.raw, t


def msm_g2(bases: bytes, scalars: bytes):
    n = len(scalars) // 32
    out = C.create_string_buffer(128)
    t = lib().ob_msm_g2(_buf(bases), _buf(scalars), n, out)
    return out.raw, t


def synthetic_bases(n: int, seed: int, g2=False) -> bytes:
    out = C.create_string_buffer(n * (128 if g2 else 64))
    (lib().ob_synthetic_bases_g2 if g2 else lib().ob_synthetic_bases_g1)(n, seed, out)
    return out.raw


def random_scalars(n: int, seed: int) -> bytes:
    out = C.create_string_buffer(max(n * 32, 1))
    lib().ob_random_scalars(n, seed, out)
    return out.raw[: n * 32]


class ProvingKey:
    def __init__(self, log_n: int, seed: int):
        self.log_n = log_n
        self.h = lib().ob_setup_synthetic(log_n, seed)

    def vk(self) -> bytes:
        out = C.create_string_buffer(576)
        lib().ob_pk_vk(self.h, out)
        return out.raw

    def fixed(self) -> dict:
        out = C.create_string_buffer(448)
        lib().ob_pk_fixed(self.h, out)
        b = out.raw
        return {"alpha_g1": b[:64], "beta_g1": b[64:128], "delta_g1": b[128:192], "beta_g2": b[192:320],
                "delta_g2": b[320:448]}

    def query(self, which: int) -> bytes:
        n = lib().ob_pk_query(self.h, which, None)
        out = C.create_string_buffer(max(n * (128 if which == 2 else 64), 1))
        lib().ob_pk_query(self.h, which, out)
        return out.raw[: n * (128 if which == 2 else 64)]

    def prove(self, witness: bytes, r: bytes, s: bytes):
        """-> (proof 256 B, intermediates 384 B, h bytes, seconds)"""
        n = 1 << self.log_n
        proof = C.create_string_buffer(256)
        inter = C.create_string_buffer(384)
        h = C.create_string_buffer(32 * (n - 1))
        t = lib().ob_groth16_prove(self.h, _buf(witness), r, s, proof, inter, h)
        return proof.raw, inter.raw, h.raw, t

    def free(self):
        if self.h:
            lib().ob_pk_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def synthetic_witness(log_n: int, w0: bytes, w1: bytes) -> bytes:
    out = C.create_string_buffer(32 << log_n)
    lib().ob_synthetic_witness(log_n, w0, w1, out)
    return out.raw
