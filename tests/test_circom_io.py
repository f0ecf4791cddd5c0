"""iden3 .r1cs / .wtns readers (libzkb.so host code) — CPU tests; the GPU proof of a
circuit that arrives through these files is in test_gpu_groth16.py."""
import struct

import pytest

import oracle_a as o
from helpers import small_circuit


@pytest.fixture(scope="module")
def cio(zkb):
    from eigen_zeth_b200 import circom_io
    return circom_io


@pytest.mark.parametrize("header_last", [False, True])
def test_r1cs_round_trip(cio, header_last):
    cs, z = small_circuit(29)
    assert o.r1cs_is_satisfied(cs, z)
    rows = list(zip(cs["A"], cs["B"], cs["C"]))
    data = cio.write_r1cs(cs["m"], 1, 0, 2, rows, header_last=header_last)
    r = cio.read_r1cs(data)
    assert (r.n_wires, r.n_pub_out, r.n_pub_in, r.n_prv_in, r.n_constraints) == (cs["m"], 1, 0, 2, 29)
    assert r.n_public == 1 and r.n_labels == cs["m"]
    assert (1 << r.log_n) >= 29 + 2 > (1 << (r.log_n - 1))
    got = list(cio.csr_to_constraints(r.matrices))
    assert got == [tuple([(i, c % o.R) for i, c in lc] for lc in row) for row in rows]
    # re-export is byte-identical
    assert cio.write_r1cs(r.n_wires, r.n_pub_out, r.n_pub_in, r.n_prv_in, got, header_last=header_last) == data


def test_wtns_round_trip(cio):
    _, z = small_circuit(29)
    data = cio.write_wtns(z)
    raw = cio.read_wtns(data)
    assert raw == b"".join(v.to_bytes(32, "little") for v in z)
    assert cio.read_wtns(cio.write_wtns(raw)) == raw


def test_empty_linear_combinations_and_zero_constraints(cio):
    # a constraint with empty A, B (0 * 0 = 0 written by circom for `signal === signal` folds) and an empty system
    r = cio.read_r1cs(cio.write_r1cs(3, 0, 1, 1, [([], [], []), ([(1, 5)], [(0, 1)], [(2, 5)])]))
    assert r.matrices[0][0] == [0, 0, 1] and r.matrices[2][1] == [2]
    r = cio.read_r1cs(cio.write_r1cs(1, 0, 0, 0, []))
    assert r.n_constraints == 0 and r.matrices[0][0] == [0]


def test_malformed_files_are_rejected(cio, zkb):
    cs, z = small_circuit(5)
    good = cio.write_r1cs(cs["m"], 1, 0, 2, list(zip(cs["A"], cs["B"], cs["C"])))
    bad_cases = {
        "magic": b"r1cx" + good[4:],
        "version": good[:4] + struct.pack("<I", 2) + good[8:],
        "truncated": good[:-7],
        "trailing bytes inside the constraint section are not ignored": None,
        "prime": good.replace(o.R.to_bytes(32, "little"), o.Q.to_bytes(32, "little"), 1),
        "wire out of range": cio.write_r1cs(2, 0, 1, 0, [([(2, 1)], [(0, 1)], [(1, 1)])]),
        "more public/private signals than wires": cio.write_r1cs(2, 1, 1, 1, []),
    }
    for name, data in bad_cases.items():
        if data is None:
            continue
        with pytest.raises(zkb.ZkbError):
            cio.read_r1cs(data)
    # a coefficient equal to r is not canonical
    nc = bytearray(cio.write_r1cs(2, 0, 1, 0, [([(1, 7)], [(0, 1)], [(1, 7)])]))
    at = bytes(nc).index((7).to_bytes(32, "little"))
    nc[at:at + 32] = o.R.to_bytes(32, "little")
    with pytest.raises(zkb.ZkbError):
        cio.read_r1cs(bytes(nc))
    # duplicate header section
    hdr_at = 12
    size = struct.unpack_from("<Q", good, hdr_at + 4)[0]
    dup = good[:8] + struct.pack("<I", 4) + good[12:] + good[hdr_at:hdr_at + 12 + size]
    with pytest.raises(zkb.ZkbError):
        cio.read_r1cs(dup)
    # witness: wire 0 must be 1, values canonical, section size must match the count
    w = cio.write_wtns(z)
    with pytest.raises(zkb.ZkbError):
        cio.read_wtns(cio.write_wtns([2] + z[1:]))
    with pytest.raises(zkb.ZkbError):
        cio.read_wtns(cio.write_wtns(b"".join(v.to_bytes(32, "little") for v in z[:-1]) + o.R.to_bytes(32, "little")))
    with pytest.raises(zkb.ZkbError):
        cio.read_wtns(w[:-32])
    with pytest.raises(zkb.ZkbError):
        cio.read_wtns(b"r1cs" + w[4:])
    # unknown section types are skipped (snarkjs adds custom-gate sections 4 and 5 to newer files)
    extra = good[:8] + struct.pack("<I", 4) + good[12:] + struct.pack("<IQ", 4, 3) + b"abc"
    assert cio.read_r1cs(extra).n_constraints == 5


def test_parsers_survive_mutated_files(cio, zkb):
    """truncations, byte flips, insertions and hostile 64-bit lengths: accepted or ZKB_EINVAL, never a crash."""
    import random
    cs, z = small_circuit(9)
    files = [cio.write_r1cs(cs["m"], 1, 0, 2, list(zip(cs["A"], cs["B"], cs["C"]))), cio.write_wtns(z)]
    rng = random.Random(1)
    rejected = 0
    for it in range(4000):
        b = bytearray(files[it & 1])
        kind = rng.randrange(4)
        if kind == 0:
            b = b[:rng.randrange(len(b) + 1)]
        elif kind == 1:
            for _ in range(rng.randrange(1, 4)):
                b[rng.randrange(len(b))] = rng.randrange(256)
        elif kind == 2:
            i = rng.randrange(len(b))
            b[i:i] = bytes(rng.randrange(256) for _ in range(rng.randrange(1, 9)))
        else:
            i = rng.randrange(len(b) - 8)
            b[i:i + 8] = rng.choice([0, 2**32 - 1, 2**64 - 1, 2**63, len(b), len(b) - i]).to_bytes(8, "little")
        try:
            (cio.read_wtns if it & 1 else cio.read_r1cs)(bytes(b))
        except zkb.ZkbError:
            rejected += 1
    assert rejected > 1000


# ---------------------------------------------------------------------------- .zkey (snarkjs Groth16 key)
@pytest.fixture(scope="module")
def zkey_case():
    """a tiny circuit, its Oracle A setup, and the .zkey bytes oracle_zkey writes for it"""
    import oracle_zkey as zkm
    cs, z = small_circuit(5)
    rng = o.Xoshiro256ss(0x2E7)
    tau, alpha, beta, gamma, delta, r, s = [rng.fr() or 1 for _ in range(7)]
    pk, vk = o.groth16_setup(cs, tau, alpha, beta, gamma, delta)
    data = zkm.zkey_bytes(cs, pk, vk, zkm.h_points_snarkjs(tau, delta, pk["n"]))
    return cs, z, pk, vk, r, s, data


def test_zkey_snarkjs_witness_map_gives_the_ark_proof(zkey_case):
    """the oracle pins the maths the GPU path relies on: the H section L_{2i+1}(tau)/delta with
    (ab - c) on the odd points of the size-2n domain yields the SAME proof bytes as ark-groth16's
    tau^i Z(tau)/delta with h = (ab - c)/Z, and the proof verifies under the key's own vk"""
    import oracle_zkey as zkm
    cs, z, pk, vk, r, s, data = zkey_case
    zk = zkm.zkey_parse(data)
    p_ark, i_ark = o.groth16_prove(pk, cs, z, r, s)
    p_zk, i_zk = zkm.groth16_prove_zkey(zk, z, r, s)
    assert p_zk == p_ark and i_zk["msm_h"] == i_ark["msm_h"]
    assert i_zk["a_evals"] == i_ark["a_evals"] and i_zk["b_evals"] == i_ark["b_evals"] and i_zk["c_evals"] == i_ark["c_evals"]
    assert o.groth16_verify(zkm.vk_of(zk), p_zk, [z[1]])
    assert not o.groth16_verify(zkm.vk_of(zk), p_zk, [(z[1] + 1) % o.R])


def test_zkey_reader_matches_the_oracle_parser(cio, zkey_case):
    import oracle_zkey as zkm
    cs, z, pk, vk, r, s, data = zkey_case
    zk, ref = cio.read_zkey(data), zkm.zkey_parse(data)
    n_cons = len(cs["A"])
    assert (zk.n_vars, zk.n_public, 1 << zk.log_n, zk.n_rows) == (cs["m"], 1, pk["n"], n_cons + 2)
    g1s = lambda pts: b"".join(o.g1_to_bytes(P) for P in pts)
    assert zk.queries["a_query"] == g1s(ref["a_query"]) == g1s(pk["a_query"])
    assert zk.queries["b_g1_query"] == g1s(ref["b_g1_query"])
    assert zk.queries["b_g2_query"] == b"".join(o.g2_to_bytes(P) for P in ref["b_g2_query"])
    assert zk.queries["l_query"] == g1s(ref["c_query"]) == g1s(pk["l_query"])
    assert zk.queries["h_query"] == g1s(ref["h_query"]) and len(zk.queries["h_query"]) == 64 * pk["n"]
    for name in ("alpha_g1", "beta_g1", "delta_g1"):
        assert zk.fixed[name] == o.g1_to_bytes(ref[name])
    for name in ("beta_g2", "gamma_g2", "delta_g2"):
        assert zk.fixed[name] == o.g2_to_bytes(ref[name])
    assert zk.ic == g1s(vk["gamma_abc_g1"])
    assert zk.vk_bytes() == (o.g1_to_bytes(vk["alpha_g1"]) + o.g2_to_bytes(vk["beta_g2"]) + o.g2_to_bytes(vk["gamma_g2"])
                             + o.g2_to_bytes(vk["delta_g2"]) + g1s(vk["gamma_abc_g1"]))
    # CSR of A (with the input rows) and B == the coefficient list, whatever order the file lists it in
    want = [{}, {}]
    for mat, row, sig, v in ref["coeffs"]:
        want[mat].setdefault(row, []).append((sig, v))
    for k in range(2):
        rp, col, co = zk.matrices[k]
        got = {}
        for row in range(zk.n_rows):
            for t in range(rp[row], rp[row + 1]):
                got.setdefault(row, []).append((col[t], int.from_bytes(co[32 * t:32 * t + 32], "little")))
        assert got == want[k]
    assert want[0][n_cons] == [(0, 1)] and want[0][n_cons + 1] == [(1, 1)]


def test_zkey_entries_in_any_order_and_malformed_files(cio, zkb, zkey_case):
    import oracle_zkey as zkm
    cs, z, pk, vk, r, s, data = zkey_case
    zk = cio.read_zkey(data)
    ref = zkm.zkey_parse(data)
    # the same key with the coefficient entries reversed reads to the same matrices up to the order inside a row
    g1s = lambda pts: b"".join(o.g1_to_bytes(P) for P in pts)
    fixed = dict(zk.fixed)
    rev = zkm.zkey_bytes_raw(zk.n_vars, zk.n_public, 1 << zk.log_n, ref["coeffs"][::-1], fixed, zk.ic,
                             zk.queries["a_query"], zk.queries["b_g1_query"], zk.queries["b_g2_query"],
                             zk.queries["l_query"], zk.queries["h_query"])
    zr = cio.read_zkey(rev)
    assert zr.queries == zk.queries and zr.fixed == zk.fixed and zr.ic == zk.ic
    for k in range(2):
        assert zr.matrices[k][0] == zk.matrices[k][0]
        rows = lambda mz: [sorted((mz[1][t], mz[2][32 * t:32 * t + 32]) for t in range(mz[0][j], mz[0][j + 1]))
                           for j in range(zk.n_rows)]
        assert rows(zr.matrices[k]) == rows(zk.matrices[k])

    def bad(mutated):
        with pytest.raises(zkb.ZkbError):
            cio.read_zkey(bytes(mutated))

    bad(b"zkez" + data[4:])                                      # magic
    bad(data[:4] + struct.pack("<I", 2) + data[8:])              # version
    bad(data[:-1])                                               # truncated last section
    bad(data[:len(data) // 2])
    i = data.index(struct.pack("<IQ", 1, 4)) + 12
    bad(data[:i] + struct.pack("<I", 2) + data[i + 4:])          # protocol 2 = plonk
    h = data.index(o.Q.to_bytes(32, "little"))
    bad(data[:h] + o.R.to_bytes(32, "little") + data[h + 32:])   # another base field
    # a coordinate that is not reduced (first point of section 5 := q)
    s5 = data.index(struct.pack("<IQ", 5, 64 * zk.n_vars)) + 12
    bad(data[:s5] + o.Q.to_bytes(32, "little") + data[s5 + 32:])
    # coefficient entry with a signal / row / matrix out of range, and a value >= r
    s4 = data.index(struct.pack("<IQ", 4, 4 + 44 * len(ref["coeffs"]))) + 12
    bad(data[:s4 + 4] + struct.pack("<I", 2) + data[s4 + 8:])
    bad(data[:s4 + 8] + struct.pack("<I", 1 << zk.log_n) + data[s4 + 12:])
    bad(data[:s4 + 12] + struct.pack("<I", zk.n_vars) + data[s4 + 16:])
    bad(data[:s4 + 16] + o.R.to_bytes(32, "little") + data[s4 + 48:])
    bad(data[:s4] + struct.pack("<I", len(ref["coeffs"]) + 1) + data[s4 + 4:])   # count disagrees with the section size


def test_zkey_with_three_public_inputs_through_reader_and_verify_ex(cio, zkb):
    """the key-side bookkeeping for n_public != 1: four input rows in A, four IC points, a C section
    without the instance variables; the oracle's snarkjs-style proof verifies through zkb_groth16_verify_ex"""
    import oracle_zkey as zkm
    cs, z = small_circuit(5)
    cs["n_pub"] = 3
    rng = o.Xoshiro256ss(77)
    tau, alpha, beta, gamma, delta, r, s = [rng.fr() or 1 for _ in range(7)]
    pk, vk = o.groth16_setup(cs, tau, alpha, beta, gamma, delta)
    data = zkm.zkey_bytes(cs, pk, vk, zkm.h_points_snarkjs(tau, delta, pk["n"]))
    zk, ref = cio.read_zkey(data), zkm.zkey_parse(data)
    assert (zk.n_public, zk.n_rows, len(zk.ic), len(zk.queries["l_query"])) == (3, 5 + 4, 64 * 4, 64 * (cs["m"] - 4))
    proof_pts, _ = zkm.groth16_prove_zkey(ref, z, r, s)
    assert proof_pts == o.groth16_prove(pk, cs, z, r, s)[0]
    proof = o.g1_to_bytes(proof_pts[0]) + o.g2_to_bytes(proof_pts[1]) + o.g1_to_bytes(proof_pts[2])
    inputs = b"".join(v.to_bytes(32, "little") for v in z[1:4])
    assert zkb.groth16_verify_ex(zk.vk_fixed(), zk.ic, proof, inputs)
    assert not zkb.groth16_verify_ex(zk.vk_fixed(), zk.ic, proof, inputs[32:] + inputs[:32])


def test_zkey_parser_survives_mutated_files(cio, zkb, zkey_case):
    """as for .r1cs / .wtns: truncations, byte flips, insertions, hostile lengths and counts end in an
    accepted key or ZKB_EINVAL, never in a crash or an out-of-bounds write (read_zkey sizes every
    buffer from zkb_zkey_inspect on the same bytes)"""
    import random
    data = zkey_case[6]
    rng = random.Random(2)
    rejected = accepted = 0
    for it in range(3000):
        b = bytearray(data)
        kind = rng.randrange(5)
        if kind == 0:
            b = b[:rng.randrange(len(b) + 1)]
        elif kind == 1:
            for _ in range(rng.randrange(1, 4)):
                b[rng.randrange(len(b))] = rng.randrange(256)
        elif kind == 2:
            i = rng.randrange(len(b))
            b[i:i] = bytes(rng.randrange(256) for _ in range(rng.randrange(1, 9)))
        elif kind == 3:
            i = rng.randrange(len(b) - 8)
            b[i:i + 8] = rng.choice([0, 2**32 - 1, 2**64 - 1, 2**63, len(b), len(b) - i]).to_bytes(8, "little")
        else:                                   # a 32-bit field (count, row, signal, n_vars, domain ...) goes wild
            i = rng.randrange(len(b) - 4)
            b[i:i + 4] = rng.choice([0, 1, 2**31, 2**32 - 1, 2**16]).to_bytes(4, "little")
        try:
            cio.read_zkey(bytes(b))
            accepted += 1
        except zkb.ZkbError:
            rejected += 1
    assert rejected > 1500 and accepted > 0
