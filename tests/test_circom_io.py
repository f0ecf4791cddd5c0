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
