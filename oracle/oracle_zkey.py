"""TEST INFRASTRUCTURE ONLY (like everything under oracle/): the snarkjs `.zkey` container for Groth16
keys, restated in pure Python — a writer that makes fixtures from an Oracle A setup, a reader, and
the snarkjs witness map / prover on top of Oracle A's arithmetic.

The reference (eigen-zeth) holds no circuit, key or prover code: its prover is an external service
(src/prover/provider.rs:606), and the keys that service consumes are snarkjs / rapidsnark `.zkey`
files.  snarkjs is not in this image, so the layout below follows its
published format from memory and is NOT pinned against a file snarkjs wrote ("parity
unpinned" for this row); what IS pinned here: a key written in this layout, read back and proved
with the snarkjs witness map gives byte for byte the proof of the ark-groth16 style prover
(oracle_a.groth16_prove) for the same (r, s), and it verifies.

  file       := "zkey" version:u32(1) n_sections:u32 section*      section := id:u32 size:u64 payload
  1 header   := protocol:u32 (1 = groth16)
  2 groth16  := n8q:u32 q[n8q] n8r:u32 r[n8r] n_vars:u32 n_public:u32 domain_size:u32
                alpha_g1 beta_g1 beta_g2 gamma_g2 delta_g1 delta_g2
  3 IC       := (n_public + 1) G1         gamma_abc of the verifying key
  4 coeffs   := count:u32 (matrix:u32 row:u32 signal:u32 value[n8r])*   matrix 0 = A, 1 = B; C is not
                stored (the prover takes c = a * b on the domain); A carries one extra row per
                instance variable (row n_constraints + i holds 1 * z_i); value = coefficient * R^2
                mod r as a little-endian integer ("double Montgomery": the prover multiplies it by
                the raw witness value with ONE Montgomery product and lands in Montgomery form)
  5 A 6 B1 7 B2 8 C 9 H := points: n_vars G1, n_vars G1, n_vars G2, (n_vars - n_public - 1) G1,
                domain_size G1
  10 contributions (ignored by a prover)
  points     := affine coordinates, each a little-endian integer in MONTGOMERY form (x * 2^256 mod q);
                G2 coordinates c0 then c1; all-zero = infinity

H differs from ark-groth16's h_query (tau^i Z(tau) / delta, i < n - 1): snarkjs evaluates A, B, C on
the ODD points of the size-2n domain (coset shift w_2n), takes a*b - c there WITHOUT dividing by Z,
and its H section holds L^(2n)_{2i+1}(tau) / delta, the Lagrange basis of the size-2n domain at the
odd points: (AB - C)(x) has degree <= 2n - 2 and vanishes on the even points, so
sum_i (ab - c)(w_2n^{2i+1}) L_{2i+1}(tau) = (AB - C)(tau) = H(tau) Z(tau).  Same group element, so the
proof bytes equal the ark-style ones.
"""
from __future__ import annotations

import struct

import oracle_a as o

R_MONT = 1 << 256


def _fq_lem(x: int) -> bytes:
    return (x * R_MONT % o.Q).to_bytes(32, "little")


def h_scalars_snarkjs(tau: int, delta: int, n: int):
    """L^(2n)_{2i+1}(tau) / delta for i < n (test-only: needs the toxic waste)"""
    L2 = o.lagrange_at(tau, 2 * n)
    di = o.inv(delta, o.R)
    return [L2[2 * i + 1] * di % o.R for i in range(n)]


def h_points_snarkjs(tau: int, delta: int, n: int):
    return [o.g1_mul(o.G1_GEN, k) for k in h_scalars_snarkjs(tau, delta, n)]


def _lem_of_canonical(b: bytes) -> bytes:
    """canonical little-endian coordinates (32 bytes each) -> Montgomery form; all-zero points stay all-zero"""
    return b"".join(_fq_lem(int.from_bytes(b[i:i + 32], "little")) for i in range(0, len(b), 32))


def coeffs_of(cs):
    """the coefficient list of a .zkey for an R1CS in oracle A's form: per constraint A's terms then
    B's (snarkjs order), then the n_pub + 1 input rows of A"""
    n_cons = len(cs["A"])
    out = []
    for j in range(n_cons):
        for mat, key in ((0, "A"), (1, "B")):
            for sig, c in cs[key][j]:
                out.append((mat, j, sig, c % o.R))
    for i in range(cs["n_pub"] + 1):
        out.append((0, n_cons + i, i, 1))
    return out


def zkey_bytes_raw(n_vars, n_public, n, coeffs, fixed, ic, a, b1, b2, c, h, contributions: bytes = b"") -> bytes:
    """all points as canonical affine bytes (fixed: dict alpha_g1, beta_g1, beta_g2, gamma_g2, delta_g1,
    delta_g2); coeffs: [(matrix, row, signal, value int)]"""
    sections = [(1, struct.pack("<I", 1))]
    hdr = struct.pack("<I", 32) + o.Q.to_bytes(32, "little") + struct.pack("<I", 32) + o.R.to_bytes(32, "little")
    hdr += struct.pack("<III", n_vars, n_public, n)
    for name in ("alpha_g1", "beta_g1", "beta_g2", "gamma_g2", "delta_g1", "delta_g2"):
        hdr += _lem_of_canonical(fixed[name])
    sections.append((2, hdr))
    sections.append((3, _lem_of_canonical(ic)))
    body = bytearray(struct.pack("<I", len(coeffs)))
    for mat, row, sig, v in coeffs:
        body += struct.pack("<III", mat, row, sig) + (v * R_MONT * R_MONT % o.R).to_bytes(32, "little")
    sections.append((4, bytes(body)))
    for sid, pts in ((5, a), (6, b1), (7, b2), (8, c), (9, h)):
        sections.append((sid, _lem_of_canonical(pts)))
    sections.append((10, contributions or (bytes(64) + struct.pack("<I", 0))))
    out = b"zkey" + struct.pack("<II", 1, len(sections))
    for sid, payload in sections:
        out += struct.pack("<IQ", sid, len(payload)) + payload
    return out


def zkey_bytes(cs, pk, vk, h_points, contributions: bytes = b"") -> bytes:
    """a .zkey for the R1CS `cs` (oracle_a format) from an Oracle A setup (pk, vk) and the H section"""
    m, n_pub, n = cs["m"], cs["n_pub"], pk["n"]
    assert len(h_points) == n and len(pk["l_query"]) == m - n_pub - 1
    g1s = lambda pts: b"".join(o.g1_to_bytes(P) for P in pts)
    g2s = lambda pts: b"".join(o.g2_to_bytes(P) for P in pts)
    fixed = {"alpha_g1": o.g1_to_bytes(pk["alpha_g1"]), "beta_g1": o.g1_to_bytes(pk["beta_g1"]),
             "beta_g2": o.g2_to_bytes(pk["beta_g2"]), "gamma_g2": o.g2_to_bytes(vk["gamma_g2"]),
             "delta_g1": o.g1_to_bytes(pk["delta_g1"]), "delta_g2": o.g2_to_bytes(pk["delta_g2"])}
    return zkey_bytes_raw(m, n_pub, n, coeffs_of(cs), fixed, g1s(vk["gamma_abc_g1"]), g1s(pk["a_query"]),
                          g1s(pk["b_g1_query"]), g2s(pk["b_g2_query"]), g1s(pk["l_query"]), g1s(h_points), contributions)


def _fq_from_lem(b: bytes) -> int:
    v = int.from_bytes(b, "little")
    if v >= o.Q:
        raise ValueError("coordinate not reduced")
    return v * o.inv(R_MONT % o.Q, o.Q) % o.Q


def _g1_from_lem(b: bytes):
    if b == bytes(64):
        return None
    P = (_fq_from_lem(b[:32]), _fq_from_lem(b[32:]))
    if not o.g1_is_on_curve(P):
        raise ValueError("G1 point off the curve")
    return P


def _g2_from_lem(b: bytes):
    if b == bytes(128):
        return None
    c = [_fq_from_lem(b[32 * i:32 * i + 32]) for i in range(4)]
    P = ((c[0], c[1]), (c[2], c[3]))
    if not o.g2_is_on_curve(P):
        raise ValueError("G2 point off the twist")
    return P


def zkey_parse(data: bytes):
    """-> dict with the header fields, vk points, IC, coefficient list [(matrix, row, signal, value)]
    (values canonical) and the five point sections"""
    if data[:4] != b"zkey":
        raise ValueError("not a zkey")
    version, n_sections = struct.unpack_from("<II", data, 4)
    if version != 1:
        raise ValueError("unsupported zkey version")
    pos, sec = 12, {}
    for _ in range(n_sections):
        sid, size = struct.unpack_from("<IQ", data, pos)
        pos += 12
        if pos + size > len(data) or sid in sec:
            raise ValueError("truncated or duplicate section")
        sec[sid] = data[pos:pos + size]
        pos += size
    if struct.unpack_from("<I", sec[1], 0)[0] != 1:
        raise ValueError("not a groth16 key")
    h = sec[2]
    n8q = struct.unpack_from("<I", h, 0)[0]
    q = int.from_bytes(h[4:4 + n8q], "little")
    n8r = struct.unpack_from("<I", h, 4 + n8q)[0]
    r = int.from_bytes(h[8 + n8q:8 + n8q + n8r], "little")
    if (n8q, q, n8r, r) != (32, o.Q, 32, o.R):
        raise ValueError("not a BN254 key")
    p = 8 + n8q + n8r
    n_vars, n_public, domain = struct.unpack_from("<III", h, p)
    p += 12
    out = {"n_vars": n_vars, "n_public": n_public, "n": domain}
    for name, g2 in (("alpha_g1", 0), ("beta_g1", 0), ("beta_g2", 1), ("gamma_g2", 1), ("delta_g1", 0), ("delta_g2", 1)):
        size = 128 if g2 else 64
        out[name] = (_g2_from_lem if g2 else _g1_from_lem)(h[p:p + size])
        p += size
    if p != len(h):
        raise ValueError("groth16 header has trailing bytes")
    pts1 = lambda b: [_g1_from_lem(b[i:i + 64]) for i in range(0, len(b), 64)]
    pts2 = lambda b: [_g2_from_lem(b[i:i + 128]) for i in range(0, len(b), 128)]
    out["ic"] = pts1(sec[3])
    count = struct.unpack_from("<I", sec[4], 0)[0]
    if len(sec[4]) != 4 + 44 * count:
        raise ValueError("coefficient section size")
    rinv2 = pow(o.inv(R_MONT % o.R, o.R), 2, o.R)
    coeffs = []
    for i in range(count):
        mat, row, sig = struct.unpack_from("<III", sec[4], 4 + 44 * i)
        v = int.from_bytes(sec[4][16 + 44 * i:48 + 44 * i], "little")
        if mat > 1 or row >= domain or sig >= n_vars or v >= o.R:
            raise ValueError("bad coefficient entry")
        coeffs.append((mat, row, sig, v * rinv2 % o.R))
    out["coeffs"] = coeffs
    out["a_query"], out["b_g1_query"], out["b_g2_query"] = pts1(sec[5]), pts1(sec[6]), pts2(sec[7])
    out["c_query"], out["h_query"] = pts1(sec[8]), pts1(sec[9])
    if (len(out["ic"]), len(out["a_query"]), len(out["b_g1_query"]), len(out["b_g2_query"]), len(out["c_query"]),
            len(out["h_query"])) != (n_public + 1, n_vars, n_vars, n_vars, n_vars - n_public - 1, domain):
        raise ValueError("section sizes do not match the header")
    return out


def vk_of(zk):
    return {"alpha_g1": zk["alpha_g1"], "beta_g2": zk["beta_g2"], "gamma_g2": zk["gamma_g2"],
            "delta_g2": zk["delta_g2"], "gamma_abc_g1": zk["ic"]}


def witness_map_snarkjs(zk, z):
    """(a*b - c) on the odd points of the size-2n domain: the scalars of the H-section MSM
    (snarkjs groth16_prove: buildABC1, ifft, batchApplyKey(w_2n), fft, joinABC)"""
    n = zk["n"]
    a, b = [0] * n, [0] * n
    for mat, row, sig, v in zk["coeffs"]:
        t = a if mat == 0 else b
        t[row] = (t[row] + v * z[sig]) % o.R
    c = [x * y % o.R for x, y in zip(a, b)]
    w2n = o.fr_root_of_unity(n.bit_length())        # log2(2n) = bit_length(n)
    ev = [o.coset_ntt(o.ntt(v, inverse=True), shift=w2n) for v in (a, b, c)]
    return [(x * y - t) % o.R for x, y, t in zip(*ev)], (a, b, c)


def groth16_prove_zkey(zk, z, r, s):
    n_inst = zk["n_public"] + 1
    h, evals = witness_map_snarkjs(zk, z)
    msm_a = o.msm_g1(zk["a_query"], z)
    msm_b1 = o.msm_g1(zk["b_g1_query"], z)
    msm_b2 = o.msm_g2(zk["b_g2_query"], z)
    msm_l = o.msm_g1(zk["c_query"], z[n_inst:])
    msm_h = o.msm_g1(zk["h_query"], h)
    A = o.g1_add(o.g1_add(zk["alpha_g1"], msm_a), o.g1_mul(zk["delta_g1"], r))
    B1 = o.g1_add(o.g1_add(zk["beta_g1"], msm_b1), o.g1_mul(zk["delta_g1"], s))
    B2 = o.g2_add(o.g2_add(zk["beta_g2"], msm_b2), o.g2_mul(zk["delta_g2"], s))
    C = o.g1_add(o.g1_add(msm_l, msm_h), o.g1_add(o.g1_mul(A, s), o.g1_mul(B1, r)))
    C = o.g1_add(C, o.g1_neg(o.g1_mul(zk["delta_g1"], r * s % o.R)))
    inter = {"h": h, "a_evals": evals[0], "b_evals": evals[1], "c_evals": evals[2],
             "msm_a": msm_a, "msm_b1": msm_b1, "msm_b2": msm_b2, "msm_l": msm_l, "msm_h": msm_h}
    return (A, B2, C), inter
