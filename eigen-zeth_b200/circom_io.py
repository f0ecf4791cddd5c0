"""circom / snarkjs circuit files -> zkb_pk_load / zkb_groth16_prove inputs.

The reference never sees a circuit: its prover is an external service (src/prover/provider.rs:606)
and the tree holds no R1CS, witness or key.  A deployment that swaps that service for this backend
brings its circuit as the iden3 binary files circom writes:
  * `.r1cs`  — constraint system  -> the three CSR matrices of `zkb_pk_desc`;
  * `.wtns`  — witness            -> the `witness` argument of `zkb_groth16_prove`;
  * `.zkey`  — snarkjs / rapidsnark Groth16 proving key -> a complete `zkb_pk_desc` with
    `h_basis = 1` (matrices A and B, the five queries, the fixed points) plus the verifying key.
Reading is done by libzkb.so (csrc/circom_io.cpp: zkb_r1cs_inspect / zkb_r1cs_read /
zkb_wtns_inspect / zkb_wtns_read / zkb_zkey_inspect / zkb_zkey_read) so a Rust or C++ host gets the same parser; this module is the
ctypes face of it plus writers (export of the synthetic benchmark circuit, test fixtures).
"""
from __future__ import annotations

import ctypes as C
import struct
from dataclasses import dataclass

R = 21888242871839275222246405745257275088548364400416034343698204186575808495617


@dataclass
class R1cs:
    n_wires: int
    n_pub_out: int
    n_pub_in: int
    n_prv_in: int
    n_labels: int
    n_constraints: int
    matrices: list          # three (row_ptr list[int], col list[int], coeff bytes): A, B, C

    @property
    def n_public(self) -> int:
        """instance variables after the constant 1 (zkb_pk_desc.n_public)."""
        return self.n_pub_out + self.n_pub_in

    @property
    def log_n(self) -> int:
        """smallest ark-groth16 domain: n_constraints + n_public + 1 <= 2^log_n."""
        return max(1, (self.n_constraints + self.n_public + 1 - 1).bit_length())


def read_r1cs(data: bytes) -> R1cs:
    from . import lib as zkb
    lib = zkb.load_library()
    info = zkb.R1csInfo()
    rc = lib.zkb_r1cs_inspect(data, len(data), C.byref(info))
    if rc:
        raise zkb.ZkbError(rc, "not a BN254 .r1cs file (magic, version, prime, truncation or a non-canonical term)")
    m = info.n_constraints
    row_ptr = [(C.c_uint32 * (m + 1))() for _ in range(3)]
    col = [(C.c_uint32 * max(1, info.nnz[k]))() for k in range(3)]
    coeff = [C.create_string_buffer(max(1, 32 * info.nnz[k])) for k in range(3)]
    ptrs = lambda arrs: (C.c_void_p * 3)(*[C.cast(a, C.c_void_p) for a in arrs])
    rc = lib.zkb_r1cs_read(data, len(data), ptrs(row_ptr), ptrs(col), ptrs(coeff))
    if rc:
        raise zkb.ZkbError(rc, "zkb_r1cs_read failed")
    mats = [(list(row_ptr[k]), list(col[k])[:info.nnz[k]], coeff[k].raw[:32 * info.nnz[k]]) for k in range(3)]
    return R1cs(info.n_wires, info.n_pub_out, info.n_pub_in, info.n_prv_in, info.n_labels, m, mats)


def read_wtns(data: bytes) -> bytes:
    """-> n_witness * 32 bytes, canonical little-endian, wire 0 = 1."""
    from . import lib as zkb
    lib = zkb.load_library()
    n = C.c_uint32(0)
    rc = lib.zkb_wtns_inspect(data, len(data), C.byref(n))
    if rc:
        raise zkb.ZkbError(rc, "not a BN254 .wtns file")
    out = C.create_string_buffer(max(1, 32 * n.value))
    rc = lib.zkb_wtns_read(data, len(data), out)
    if rc:
        raise zkb.ZkbError(rc, "witness holds a non-canonical value or wire 0 is not 1")
    return out.raw[:32 * n.value]


@dataclass
class Zkey:
    n_vars: int
    n_public: int
    log_n: int
    n_rows: int             # rows of A / B present in the key (constraints + the n_public + 1 input rows)
    matrices: list          # two (row_ptr list[int], col list[int], coeff bytes): A, B
    fixed: dict             # alpha_g1, beta_g1, delta_g1 (64 B), beta_g2, gamma_g2, delta_g2 (128 B)
    ic: bytes               # 64 * (n_public + 1)
    queries: dict           # a_query, b_g1_query, b_g2_query, l_query (the key's C section), h_query

    def vk_fixed(self) -> bytes:
        """alpha_g1 | beta_g2 | gamma_g2 | delta_g2: the first argument of zkb_groth16_verify_ex (ic is the second)"""
        f = self.fixed
        return f["alpha_g1"] + f["beta_g2"] + f["gamma_g2"] + f["delta_g2"]

    def vk_bytes(self) -> bytes:
        """the verifying key in the library's layout (one public input: ZKB_VK_BYTES = 576)"""
        return self.vk_fixed() + self.ic


def read_zkey(data: bytes) -> Zkey:
    from . import lib as zkb
    lib = zkb.load_library()
    info = zkb.ZkeyInfo()
    rc = lib.zkb_zkey_inspect(data, len(data), C.byref(info))
    if rc:
        raise zkb.ZkbError(rc, "not a BN254 Groth16 .zkey (magic, version, protocol, primes, section sizes or an index out of range)")
    m, npub, n = info.n_vars, info.n_public, 1 << info.log_n
    d = zkb.ZkeyData()
    keep = {}

    def buf(name, size):
        keep[name] = C.create_string_buffer(max(1, size))
        return C.cast(keep[name], C.c_void_p)

    row_ptr = [(C.c_uint32 * (info.n_rows + 1))() for _ in range(2)]
    col = [(C.c_uint32 * max(1, info.nnz[k]))() for k in range(2)]
    for k in range(2):
        d.row_ptr[k] = C.cast(row_ptr[k], C.c_void_p)
        d.col[k] = C.cast(col[k], C.c_void_p)
        d.coeff[k] = buf("coeff%d" % k, 32 * info.nnz[k])
    sizes = {"alpha_g1": 64, "beta_g1": 64, "delta_g1": 64, "beta_g2": 128, "gamma_g2": 128, "delta_g2": 128,
             "ic": 64 * (npub + 1), "a_query": 64 * m, "b_g1_query": 64 * m, "b_g2_query": 128 * m,
             "c_query": 64 * (m - npub - 1), "h_query": 64 * n}
    for name, size in sizes.items():
        setattr(d, name, buf(name, size))
    rc = lib.zkb_zkey_read(data, len(data), C.byref(d))
    if rc:
        raise zkb.ZkbError(rc, "zkb_zkey_read failed (a coordinate or coefficient is not reduced)")
    raw = lambda name: keep[name].raw[:sizes[name]]
    mats = [(list(row_ptr[k]), list(col[k])[:info.nnz[k]], keep["coeff%d" % k].raw[:32 * info.nnz[k]]) for k in range(2)]
    fixed = {k: raw(k) for k in ("alpha_g1", "beta_g1", "delta_g1", "beta_g2", "gamma_g2", "delta_g2")}
    queries = {"a_query": raw("a_query"), "b_g1_query": raw("b_g1_query"), "b_g2_query": raw("b_g2_query"),
               "l_query": raw("c_query"), "h_query": raw("h_query")}
    return Zkey(m, npub, info.log_n, info.n_rows, mats, fixed, raw("ic"), queries)


def load_zkey(ctx, data: bytes):
    """.zkey bytes -> a resident ProvingKey on `ctx` (zkb_pk_load with h_basis = 1)"""
    zk = read_zkey(data)
    fixed = {k: zk.fixed[k] for k in ("alpha_g1", "beta_g1", "delta_g1", "beta_g2", "delta_g2")}
    return ctx.pk_load(zk.log_n, zk.n_vars, zk.n_public, zk.matrices, zk.queries, fixed, vk=zk.vk_bytes(), h_basis=1), zk


# ---- writers -------------------------------------------------------------------------------------
def _section(kind: int, payload: bytes) -> bytes:
    return struct.pack("<IQ", kind, len(payload)) + payload


def write_r1cs(n_wires: int, n_pub_out: int, n_pub_in: int, n_prv_in: int, constraints, n_labels=None,
               header_last: bool = False) -> bytes:
    """constraints: iterable of (A, B, C), each a list of (wire, coefficient int).  `header_last`
    writes the sections in the order 2, 3, 1 (the format allows any order)."""
    body = bytearray()
    count = 0
    for abc in constraints:
        for lc in abc:
            body += struct.pack("<I", len(lc))
            for wire, coeff in lc:
                body += struct.pack("<I", wire) + int(coeff % R).to_bytes(32, "little")
        count += 1
    header = struct.pack("<I", 32) + R.to_bytes(32, "little") + struct.pack(
        "<IIIIQI", n_wires, n_pub_out, n_pub_in, n_prv_in, n_wires if n_labels is None else n_labels, count)
    labels = b"".join(struct.pack("<Q", i) for i in range(n_wires))
    sections = [_section(2, bytes(body)), _section(3, labels)]
    sections = sections + [_section(1, header)] if header_last else [_section(1, header)] + sections
    return b"r1cs" + struct.pack("<II", 1, 3) + b"".join(sections)


def write_wtns(values) -> bytes:
    """values: ints (or one bytes object of 32-byte canonical elements)."""
    raw = values if isinstance(values, (bytes, bytearray)) else b"".join(int(v % R).to_bytes(32, "little") for v in values)
    header = struct.pack("<I", 32) + R.to_bytes(32, "little") + struct.pack("<I", len(raw) // 32)
    return b"wtns" + struct.pack("<II", 2, 2) + _section(1, header) + _section(2, bytes(raw))


def csr_to_constraints(matrices):
    """inverse view of R1cs.matrices (for re-export)."""
    m = len(matrices[0][0]) - 1
    for j in range(m):
        row = []
        for rp, col, coeff in matrices:
            row.append([(col[t], int.from_bytes(coeff[32 * t:32 * t + 32], "little")) for t in range(rp[j], rp[j + 1])])
        yield tuple(row)
