"""Groth16 prove on the GPU vs oracle B (bit-exact with identical r, s) + pairing verification."""
import pytest

import oracle_a as o
import oracle_b as ob
from helpers import fr_bytes, fr_list, g1_list, g2_list

pytestmark = pytest.mark.gpu


def _inputs(seed):
    rng = o.Xoshiro256ss(seed)
    return [fr_bytes([rng.fr()]) for _ in range(4)]


@pytest.mark.parametrize("log_n", [3, 4, 6, 8, 9, 11, 14])
def test_prove_bit_exact_vs_oracle(ctx, zkb, log_n):
    seed = 0x5EED0004 + log_n
    pk = ctx.pk_synthetic(log_n, seed)
    opk = ob.ProvingKey(log_n, seed)
    assert pk.vk == opk.vk()
    for which in range(5):
        got = pk.query(which).download()
        if which == 3:      # l_query is padded with the two instance slots (points at infinity)
            assert got[:128] == bytes(128)
            got = got[128:]
        assert got == opk.query(which), "query %d" % which
    w0, w1, r, s = _inputs(seed + 1)
    wit = zkb.synthetic_witness(log_n, w0, w1)
    assert wit == ob.synthetic_witness(log_n, w0, w1)
    proof, inter, h, timings = ctx.groth16_prove(pk, wit, r, s, want_intermediates=True, want_h=True)
    oproof, ointer, oh, _ = opk.prove(wit, r, s)
    assert h == oh
    assert inter == ointer
    assert proof == oproof
    assert zkb.groth16_verify(pk.vk, proof, wit[32:64])
    bad = bytearray(wit[32:64])
    bad[0] ^= 1
    assert not zkb.groth16_verify(pk.vk, proof, bytes(bad))
    # replay with the same (r, s) is bit-identical (idempotent GenFinalProof, provider.rs:245-274)
    proof2, _, _, _ = ctx.groth16_prove(pk, wit, r, s)
    assert proof2 == proof
    # different blinding -> different but still valid proof
    proof3, _, _, _ = ctx.groth16_prove(pk, wit, s, r)
    assert proof3 != proof and zkb.groth16_verify(pk.vk, proof3, wit[32:64])
    # the final assembly runs on the host by default (six serial 254-step ladders); the one-CTA device
    # kernel is the comparison path and must give the same bytes
    ctx.set("assemble_on_device", 1)
    try:
        proof4, _, _, _ = ctx.groth16_prove(pk, wit, r, s)
    finally:
        ctx.set("assemble_on_device", 0)
    assert proof4 == proof
    pk.free()
    opk.free()


def test_prove_small_vs_oracle_a(ctx, zkb):
    """one size also against the pure-Python oracle, including its independent pairing."""
    log_n, seed = 4, 99
    pk = ctx.pk_synthetic(log_n, seed)
    rng = o.Xoshiro256ss(seed)
    tw = [rng.fr() for _ in range(5)]
    cs = o.synthetic_r1cs(log_n)
    pka, vka = o.groth16_setup(cs, *tw)
    assert g1_list(pk.query(0).download()) == pka["a_query"]
    assert g2_list(pk.query(2).download()) == pka["b_g2_query"]
    w0, w1, r, s = _inputs(7)
    wit = zkb.synthetic_witness(log_n, w0, w1)
    z = fr_list(wit)
    assert z == o.synthetic_witness(log_n, fr_list(w0)[0], fr_list(w1)[0])
    proof, _, _, _ = ctx.groth16_prove(pk, wit, r, s)
    pa, _ = o.groth16_prove(pka, cs, z, fr_list(r)[0], fr_list(s)[0])
    got = (o.g1_from_bytes(proof[:64]), o.g2_from_bytes(proof[64:192]), o.g1_from_bytes(proof[192:]))
    assert got == pa
    assert o.groth16_verify(vka, got, [z[1]])


@pytest.mark.parametrize("log_n", [18, 20])
def test_prove_large_bit_exact_vs_oracle_b(ctx, zkb, log_n):
    """h, the five MSM sums and the proof bytes at 2^18 / 2^20 constraints against oracle B (a pairing
    check alone accepts any valid proof; this pins the intermediates too)."""
    seed = 0x5EED0004 + log_n
    pk = ctx.pk_synthetic(log_n, seed)
    opk = ob.ProvingKey(log_n, seed)
    assert pk.vk == opk.vk()
    w0, w1, r, s = _inputs(seed + 1)
    wit = zkb.synthetic_witness(log_n, w0, w1)
    proof, inter, h, _ = ctx.groth16_prove(pk, wit, r, s, want_intermediates=True, want_h=True)
    oproof, ointer, oh, _ = opk.prove(wit, r, s)
    assert h == oh
    assert inter == ointer
    assert proof == oproof
    assert zkb.groth16_verify(pk.vk, proof, wit[32:64])
    pk.free()
    opk.free()


@pytest.mark.parametrize("log_n", [18, 20])
def test_prove_large_pairing_verified(ctx, zkb, log_n):
    seed = 0x5EED0004 + log_n
    pk = ctx.pk_synthetic(log_n, seed)
    w0, w1, r, s = _inputs(seed)
    wit = zkb.synthetic_witness(log_n, w0, w1)
    proof, _, _, t = ctx.groth16_prove(pk, wit, r, s)
    assert zkb.groth16_verify(pk.vk, proof, wit[32:64])
    assert not zkb.groth16_verify(pk.vk, proof, wit[64:96])
    # a corrupted witness entry must not verify
    badw = bytearray(wit)
    badw[32 * 1000] ^= 1
    proof_bad, _, _, _ = ctx.groth16_prove(pk, bytes(badw), r, s)
    assert not zkb.groth16_verify(pk.vk, proof_bad, wit[32:64])
    print("prove 2^%d: %s" % (log_n, t))
    pk.free()


@pytest.mark.parametrize("log_n,world", [(6, 2), (10, 3), (14, 4)])
def test_sharded_prove_equals_single_gpu(ctx, zkb, log_n, world):
    """point-range sharding of every MSM, emulated on one GPU: the `world` sharded
    keys produce partial sums whose slot-wise sum assembles to the same proof bytes."""
    from eigen_zeth_b200.sharding import combine_partial_sums
    seed = 0x5EED0004 + log_n
    w0, w1, r, s = _inputs(seed + 3)
    wit = zkb.synthetic_witness(log_n, w0, w1)
    pk = ctx.pk_synthetic(log_n, seed)
    want, inter, _, _ = ctx.groth16_prove(pk, wit, r, s, want_intermediates=True)
    parts = []
    for rank in range(world):
        spk = ctx.pk_synthetic_shard(log_n, seed, rank, world)
        assert spk.vk == pk.vk
        with pytest.raises(zkb.ZkbError):
            ctx.groth16_prove(spk, wit, r, s)          # a sharded key cannot prove alone
        part, _ = ctx.groth16_prove_partial(spk, wit)
        parts.append(part)
        last = spk
    sums = combine_partial_sums(parts, lambda pts, g2: ctx.point_sum(pts, g2=g2))
    assert sums == inter
    assert ctx.groth16_assemble(last, sums, r, s) == want
    assert zkb.groth16_verify(pk.vk, want, wit[32:64])


def _csr(rows):
    rp, col, coeff = [0], [], b""
    for row in rows:
        for c_, v in row:
            col.append(c_)
            coeff += int(v).to_bytes(32, "little")
        rp.append(len(col))
    return rp, col, coeff


@pytest.mark.parametrize("log_n", [4, 9, 12])
def test_pk_load_external_circuit(ctx, zkb, log_n):
    """zkb_pk_load: R1CS matrices and ark-groth16-style queries produced OUTSIDE the library (here by
    oracle B) are ingested as bytes; the proof must equal oracle B's, byte for byte."""
    seed = 0x5EED0004 + 77 + log_n
    opk = ob.ProvingKey(log_n, seed)
    cs = o.synthetic_r1cs(log_n)
    mats = [_csr(cs[k]) for k in ("A", "B", "C")]
    queries = {n: opk.query(i) for i, n in enumerate(("a_query", "b_g1_query", "b_g2_query", "l_query", "h_query"))}
    pk = ctx.pk_load(log_n, cs["m"], cs["n_pub"], mats, queries, opk.fixed(), vk=opk.vk())
    w0, w1, r, s = _inputs(seed)
    wit = ob.synthetic_witness(log_n, w0, w1)
    proof, inter, h, _ = ctx.groth16_prove(pk, wit, r, s, want_intermediates=True, want_h=True)
    oproof, ointer, oh, _ = opk.prove(wit, r, s)
    assert h == oh and inter == ointer and proof == oproof
    assert zkb.groth16_verify(opk.vk(), proof, wit[32:64])
    # error behaviour: inconsistent dimensions / out-of-range column / non-canonical coefficient
    with pytest.raises((zkb.ZkbError, ValueError)):     # the binding already rejects the short h_query
        ctx.pk_load(log_n - 1, cs["m"], cs["n_pub"], mats, queries, opk.fixed())
    short = dict(queries, a_query=queries["a_query"][:-64])
    with pytest.raises(ValueError, match="a_query"):
        ctx.pk_load(log_n, cs["m"], cs["n_pub"], mats, short, opk.fixed())
    off = bytearray(queries["a_query"])
    off[0] ^= 1                                          # x changed: the point leaves the curve
    with pytest.raises(zkb.ZkbError, match="not on the curve"):
        ctx.pk_load(log_n, cs["m"], cs["n_pub"], mats, dict(queries, a_query=bytes(off)), opk.fixed())
    with pytest.raises(ValueError, match="witness holds"):
        ctx.groth16_prove(pk, wit[:-32], r, s)
    bad = [(mats[0][0], [cs["m"]] + mats[0][1][1:], mats[0][2]), mats[1], mats[2]]
    with pytest.raises(zkb.ZkbError):
        ctx.pk_load(log_n, cs["m"], cs["n_pub"], bad, queries, opk.fixed())
    bad = [(mats[0][0], mats[0][1], int(o.R).to_bytes(32, "little") + mats[0][2][32:]), mats[1], mats[2]]
    with pytest.raises(zkb.ZkbError):
        ctx.pk_load(log_n, cs["m"], cs["n_pub"], bad, queries, opk.fixed())
    pk.free()
    opk.free()


@pytest.mark.parametrize("n_constraints", [6, 27])
def test_circuit_from_r1cs_and_wtns_files(ctx, zkb, n_constraints):
    """a circuit that is not the synthetic chain arrives as iden3 .r1cs / .wtns bytes
    (wire count and constraint count not powers of two), is parsed by libzkb.so, keyed by oracle A's
    general ark-groth16 setup and proved on the GPU; proof = oracle A's proof, byte for byte."""
    from eigen_zeth_b200 import circom_io as cio
    from helpers import small_circuit
    cs, z = small_circuit(n_constraints)
    r1 = cio.read_r1cs(cio.write_r1cs(cs["m"], 1, 0, 2, list(zip(cs["A"], cs["B"], cs["C"])), header_last=bool(n_constraints & 1)))
    wit = cio.read_wtns(cio.write_wtns(z))
    assert len(wit) == 32 * r1.n_wires
    rng = o.Xoshiro256ss(0xC1C0 + n_constraints)
    tau, alpha, beta, gamma, delta, r, s = [rng.fr() or 1 for _ in range(7)]
    opk, ovk = o.groth16_setup(cs, tau, alpha, beta, gamma, delta)
    g1s = lambda pts: b"".join(o.g1_to_bytes(P) for P in pts)
    queries = {"a_query": g1s(opk["a_query"]), "b_g1_query": g1s(opk["b_g1_query"]),
               "b_g2_query": b"".join(o.g2_to_bytes(P) for P in opk["b_g2_query"]),
               "l_query": g1s(opk["l_query"]), "h_query": g1s(opk["h_query"])}
    fixed = {"alpha_g1": o.g1_to_bytes(opk["alpha_g1"]), "beta_g1": o.g1_to_bytes(opk["beta_g1"]),
             "delta_g1": o.g1_to_bytes(opk["delta_g1"]), "beta_g2": o.g2_to_bytes(opk["beta_g2"]),
             "delta_g2": o.g2_to_bytes(opk["delta_g2"])}
    assert (1 << r1.log_n) == o.domain_size(cs)
    pk = ctx.pk_load(r1.log_n, r1.n_wires, r1.n_public, r1.matrices, queries, fixed)
    proof, _, _, _ = ctx.groth16_prove(pk, wit, r.to_bytes(32, "little"), s.to_bytes(32, "little"))
    (A, B, C), _ = o.groth16_prove(opk, cs, z, r, s)
    assert proof == o.g1_to_bytes(A) + o.g2_to_bytes(B) + o.g1_to_bytes(C)
    assert o.groth16_verify(ovk, (A, B, C), [z[1]])
    pk.free()
