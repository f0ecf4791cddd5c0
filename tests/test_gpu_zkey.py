"""a snarkjs / rapidsnark `.zkey` proving key (layout: oracle/oracle_zkey.py) is read
by libzkb.so, loaded with `h_basis = 1` and proved on the GPU with the snarkjs witness map (a*b - c
on the odd points of the size-2n domain, H section L_{2i+1}(tau)/delta).  The proof must equal the
ark-groth16 style proof of the same circuit / setup / witness / (r, s) byte for byte, so the oracles
that pin the ark path pin this one too."""
import pytest

import oracle_a as o
import oracle_b as ob
import oracle_zkey as zkm
from helpers import fr_bytes, small_circuit

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_constraints", [6, 27])
def test_zkey_small_circuit_vs_oracle_a(ctx, zkb, n_constraints):
    from eigen_zeth_b200 import circom_io as cio
    cs, z = small_circuit(n_constraints)
    rng = o.Xoshiro256ss(0x2E70 + n_constraints)
    tau, alpha, beta, gamma, delta, r, s = [rng.fr() or 1 for _ in range(7)]
    opk, ovk = o.groth16_setup(cs, tau, alpha, beta, gamma, delta)
    data = zkm.zkey_bytes(cs, opk, ovk, zkm.h_points_snarkjs(tau, delta, opk["n"]))
    pk, zk = cio.load_zkey(ctx, data)
    assert (1 << zk.log_n) == o.domain_size(cs) and zk.n_rows == n_constraints + 2
    wit = fr_bytes(z)
    proof, inter, _, _ = ctx.groth16_prove(pk, wit, fr_bytes([r]), fr_bytes([s]), want_intermediates=True)
    (A, B, C), oi = o.groth16_prove(opk, cs, z, r, s)
    assert inter == (o.g1_to_bytes(oi["msm_a"]) + o.g1_to_bytes(oi["msm_b1"]) + o.g2_to_bytes(oi["msm_b2"])
                     + o.g1_to_bytes(oi["msm_l"]) + o.g1_to_bytes(oi["msm_h"]))
    assert proof == o.g1_to_bytes(A) + o.g2_to_bytes(B) + o.g1_to_bytes(C)
    assert zkb.groth16_verify(zk.vk_bytes(), proof, wit[32:64])
    bad = bytearray(wit[32:64])
    bad[0] ^= 1
    assert not zkb.groth16_verify(zk.vk_bytes(), proof, bytes(bad))
    # a witness that violates a constraint still yields a proof, but not a valid one
    wrong = bytearray(wit)
    wrong[32 * 4] ^= 1
    p2, _, _, _ = ctx.groth16_prove(pk, bytes(wrong), fr_bytes([r]), fr_bytes([s]))
    assert not zkb.groth16_verify(zk.vk_bytes(), p2, wit[32:64])
    # the layout-specific entry points refuse such a key instead of mixing bases and scalars
    with pytest.raises(zkb.ZkbError, match="snarkjs"):
        ctx.groth16_witness_dev(pk, wit)
    pk.free()


def test_zkey_synthetic_2p10_vs_oracle_b(ctx, zkb):
    """multi-pass NTTs with the w_2n pre-scale table, 1024 variables: ark-style queries from oracle B,
    the H section from the same toxic waste, proof == oracle B's ark-style proof"""
    from eigen_zeth_b200 import circom_io as cio
    log_n, seed = 10, 0x5EED0004 + 0x2E7
    n = 1 << log_n
    opk = ob.ProvingKey(log_n, seed)
    rng = o.Xoshiro256ss(seed)
    tau, alpha, beta, gamma, delta = [rng.fr() for _ in range(5)]
    h = b"".join(ob.g1_mul_gen(k.to_bytes(32, "little")) for k in zkm.h_scalars_snarkjs(tau, delta, n))
    vk, fx = opk.vk(), opk.fixed()
    fixed = {"alpha_g1": fx["alpha_g1"], "beta_g1": fx["beta_g1"], "delta_g1": fx["delta_g1"],
             "beta_g2": fx["beta_g2"], "delta_g2": fx["delta_g2"], "gamma_g2": vk[192:320]}
    assert vk[:64] == fx["alpha_g1"] and vk[64:192] == fx["beta_g2"] and vk[320:448] == fx["delta_g2"]
    cs = o.synthetic_r1cs(log_n)
    data = zkm.zkey_bytes_raw(cs["m"], 1, n, zkm.coeffs_of(cs), fixed, vk[448:576], opk.query(0), opk.query(1),
                              opk.query(2), opk.query(3), h)
    pk, zk = cio.load_zkey(ctx, data)
    assert zk.vk_bytes() == vk and zk.log_n == log_n and zk.n_rows == n
    w0, w1, r, s = [fr_bytes([rng.fr()]) for _ in range(4)]
    wit = ob.synthetic_witness(log_n, w0, w1)
    proof, inter, _, _ = ctx.groth16_prove(pk, wit, r, s, want_intermediates=True)
    oproof, ointer, _, _ = opk.prove(wit, r, s)
    assert inter == ointer and proof == oproof
    assert zkb.groth16_verify(vk, proof, wit[32:64])
    # the same bytes again (idempotent), and through the partial entry point
    assert ctx.groth16_prove(pk, wit, r, s)[0] == proof
    assert ctx.groth16_prove_partial(pk, wit)[0] == ointer
    pk.free()
    opk.free()
