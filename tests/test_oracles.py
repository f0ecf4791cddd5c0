"""Pins of the oracles (CPU only): BN254 constants, the reference's proof
fixtures (layout, group membership), pairing bilinearity, Groth16 completeness / soundness, and
bit-for-bit agreement of Oracle B (C++) with Oracle A (pure Python)."""
import json
import os

import pytest

import oracle_a as o
import oracle_b as ob
from helpers import fr_bytes, fr_list, g1_bytes, g1_list, g2_bytes, g2_list, rand_fr

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_fixtures.json")))


def test_constants_appendix_a():
    x = 4965661367192848881
    assert o.Q == 36 * x**4 + 36 * x**3 + 24 * x**2 + 6 * x + 1
    assert o.R == 36 * x**4 + 36 * x**3 + 18 * x**2 + 6 * x + 1
    assert o.Q == 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
    assert o.R == 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
    assert (1 << 256) % o.Q == 0x0e0a77c19a07df2f666ea36f7879462c0a78eb28f5c70b3dd35d438dc58f0d9d
    assert (1 << 256) % o.R == 0x0e0a77c19a07df2f666ea36f7879462e36fc76959f60cd29ac96341c4ffffffb
    assert (-pow(o.Q, -1, 1 << 32)) % (1 << 32) == 0xe4866389
    assert (-pow(o.R, -1, 1 << 32)) % (1 << 32) == 0xefffffff
    assert (o.R - 1) % (1 << 28) == 0 and (o.R - 1) % (1 << 29) != 0
    assert pow(o.FR_ROOT_2_28, 1 << 28, o.R) == 1 and pow(o.FR_ROOT_2_28, 1 << 27, o.R) == o.R - 1
    assert pow(5, (o.R - 1) // 2, o.R) == o.R - 1          # 5 is a non-residue: a generator candidate
    assert o.g1_is_on_curve(o.G1_GEN) and o.g2_is_on_curve(o.G2_GEN) and o.g2_in_subgroup(o.G2_GEN)
    assert o.g1_mul(o.G1_GEN, o.R) is None
    assert o.ATE_LOOP == 29793968203157093288


@pytest.mark.parametrize("key", ["proof_json", "test_parse_proof_json", "worker_compact_proof_json"])
def test_reference_proof_fixtures_are_group_elements(key):
    """the reference pins layout only (Appendix C): all three fixtures parse with the parse_proof
    restatement, are on-curve / in-subgroup with the [c0, c1] reading, and off-curve when swapped."""
    A, B, C = o.parse_proof(GOLDEN[key])
    assert o.g1_is_on_curve(A) and o.g1_is_on_curve(C)
    assert o.g2_is_on_curve(B) and o.g2_in_subgroup(B)
    swapped = ((B[0][1], B[0][0]), (B[1][1], B[1][0]))
    assert not o.g2_is_on_curve(swapped)


def test_reference_parse_proof_known_answers():
    """the eight U256 values asserted by reference src/settlement/ethereum/mod.rs:515-570."""
    A, B, C = o.parse_proof(GOLDEN["test_parse_proof_json"])
    got = [A[0], A[1], B[0][0], B[0][1], B[1][0], B[1][1], C[0], C[1]]
    assert got == [int(v) for v in GOLDEN["test_parse_proof_expected"]]
    assert o.parse_public_input(GOLDEN["test_parse_public_input_json"]) == [int(GOLDEN["test_parse_public_input_expected"][0])]
    assert o.parse_public_input(GOLDEN["public_input_json"])[0] < o.R


def test_pairing_bilinear_and_nondegenerate():
    a, b = 0x1234567, 0x89abcdef
    e = o.pairing(o.G1_GEN, o.G2_GEN)
    assert e != o.F12_ONE
    assert o.pairing(o.g1_mul(o.G1_GEN, a), o.g2_mul(o.G2_GEN, b)) == o.f12_pow(e, a * b % o.R)
    assert o.f12_pow(e, o.R) == o.F12_ONE


def test_ntt_against_naive_dft():
    rng = o.Xoshiro256ss(1)
    for log_n in range(1, 8):
        v = rand_fr(rng, 1 << log_n)
        assert o.ntt(v) == o.dft_naive(v)
        assert o.ntt(v, inverse=True) == o.dft_naive(v, inverse=True)
        assert o.coset_intt(o.coset_ntt(v)) == v


def test_msm_pippenger_against_double_and_add():
    rng = o.Xoshiro256ss(2)
    pts = [o.g1_mul(o.G1_GEN, rng.fr()) for _ in range(40)] + [None]
    sc = rand_fr(rng, 41)
    sc[0], sc[1], sc[2] = 0, 1, o.R - 1
    assert o.msm_g1(pts, sc) == o.msm_g1_naive(pts, sc)
    pts2 = [o.g2_mul(o.G2_GEN, rng.fr()) for _ in range(6)]
    assert o.msm_g2(pts2, sc[:6]) == o.msm_g2_naive(pts2, sc[:6])


def test_groth16_complete_and_sound_oracle_a():
    rng = o.Xoshiro256ss(3)
    cs = o.synthetic_r1cs(3)
    z = o.synthetic_witness(3, rng.fr(), rng.fr())
    assert o.r1cs_is_satisfied(cs, z)
    pk, vk = o.groth16_setup(cs, *[rng.fr() for _ in range(5)])
    proof, inter = o.groth16_prove(pk, cs, z, rng.fr(), rng.fr())
    assert o.groth16_verify(vk, proof, [z[1]])
    assert not o.groth16_verify(vk, proof, [(z[1] + 1) % o.R])
    z_bad = list(z)
    z_bad[5] = (z_bad[5] + 1) % o.R
    assert not o.r1cs_is_satisfied(cs, z_bad)
    # proof.json round trip through the parse_proof restatement
    assert o.parse_proof(o.proof_to_json(proof)) == proof


# ------------------------------------------------------------------ Oracle B == Oracle A, bit for bit
@pytest.mark.parametrize("log_n", [1, 2, 5, 8, 11])
def test_oracle_b_ntt(log_n):
    rng = o.Xoshiro256ss(10 + log_n)
    v = rand_fr(rng, 1 << log_n)
    b = fr_bytes(v)
    assert fr_list(ob.fr_ntt(b)) == o.ntt(v)
    assert fr_list(ob.fr_ntt(b, inverse=True)) == o.ntt(v, inverse=True)
    assert fr_list(ob.fr_ntt(b, coset=True)) == o.coset_ntt(v)
    assert fr_list(ob.fr_ntt(b, inverse=True, coset=True)) == o.coset_intt(v)


def test_oracle_b_prng_bases_msm():
    n = 48
    rng = o.Xoshiro256ss(21)
    ks = [rng.fr() for _ in range(n)]
    bs = ob.synthetic_bases(n, 21)
    assert g1_list(bs) == [o.g1_mul(o.G1_GEN, k) for k in ks]
    bs2 = ob.synthetic_bases(5, 22, g2=True)
    rng2 = o.Xoshiro256ss(22)
    assert g2_list(bs2) == [o.g2_mul(o.G2_GEN, rng2.fr()) for _ in range(5)]
    rng3 = o.Xoshiro256ss(23)
    sc = fr_list(ob.random_scalars(n, 23))
    assert sc == [rng3.fr() for _ in range(n)]
    sc[0], sc[1], sc[2], sc[3] = 0, 1, o.R - 1, 1
    pts = g1_list(bs)
    pts[7] = None
    out, _ = ob.msm_g1(g1_bytes(pts), fr_bytes(sc))
    assert o.g1_from_bytes(out) == o.msm_g1_naive(pts, sc)
    out, _ = ob.msm_g2(bs2, fr_bytes(sc[:5]))
    assert o.g2_from_bytes(out) == o.msm_g2_naive(g2_list(bs2), sc[:5])


@pytest.mark.parametrize("log_n", [3, 5])
def test_oracle_b_groth16(log_n):
    seed = 300 + log_n
    pk = ob.ProvingKey(log_n, seed)
    rng = o.Xoshiro256ss(seed)
    tw = [rng.fr() for _ in range(5)]
    cs = o.synthetic_r1cs(log_n)
    pka, vka = o.groth16_setup(cs, *tw)
    assert g1_list(pk.query(0)) == pka["a_query"]
    assert g1_list(pk.query(1)) == pka["b_g1_query"]
    assert g2_list(pk.query(2)) == pka["b_g2_query"]
    assert g1_list(pk.query(3)) == pka["l_query"]
    assert g1_list(pk.query(4)) == pka["h_query"]
    vk = pk.vk()
    assert (o.g1_from_bytes(vk[:64]), o.g2_from_bytes(vk[64:192]), o.g2_from_bytes(vk[192:320]),
            o.g2_from_bytes(vk[320:448])) == (vka["alpha_g1"], vka["beta_g2"], vka["gamma_g2"], vka["delta_g2"])
    assert g1_list(vk[448:]) == vka["gamma_abc_g1"]
    w0, w1, r, s = (rng.fr() for _ in range(4))
    z = o.synthetic_witness(log_n, w0, w1)
    assert fr_list(ob.synthetic_witness(log_n, fr_bytes([w0]), fr_bytes([w1]))) == z
    proof, inter, h, _ = pk.prove(fr_bytes(z), fr_bytes([r]), fr_bytes([s]))
    pa, ia = o.groth16_prove(pka, cs, z, r, s)
    assert (o.g1_from_bytes(proof[:64]), o.g2_from_bytes(proof[64:192]), o.g1_from_bytes(proof[192:])) == pa
    assert fr_list(h) == ia["h"]
    assert g1_list(inter[:128]) == [ia["msm_a"], ia["msm_b1"]]
    assert o.g2_from_bytes(inter[128:256]) == ia["msm_b2"]
    assert g1_list(inter[256:]) == [ia["msm_l"], ia["msm_h"]]
    assert o.groth16_verify(vka, pa, [z[1]])


def test_oracle_b_closed_form_helpers():
    """the checkers bench.py and the large GPU tests rely on: sum s_i (k_i G) = (sum s_i k_i) G"""
    n = 700
    sc = ob.random_scalars(n, 5)
    bs = ob.synthetic_bases(n, 7)
    want, _ = ob.msm_g1(bs, sc)
    assert ob.msm_closed_form(sc, n, 7) == want
    vals = [int.from_bytes(sc[i * 32:i * 32 + 32], "little") for i in range(n)]
    ks = ob.random_scalars(n, 7)
    kv = [int.from_bytes(ks[i * 32:i * 32 + 32], "little") for i in range(n)]
    assert int.from_bytes(ob.fr_sum(sc, n), "little") == sum(vals) % o.R
    assert int.from_bytes(ob.fr_dot(sc, ks, n), "little") == sum(a * b for a, b in zip(vals, kv)) % o.R
    assert ob.g1_mul_gen((12345).to_bytes(32, "little")) == o.g1_to_bytes(o.g1_mul(o.G1_GEN, 12345))
    assert ob.g2_mul_gen((o.R - 1).to_bytes(32, "little")) == o.g2_to_bytes(o.g2_mul(o.G2_GEN, o.R - 1))
    b2 = ob.synthetic_bases(40, 9, g2=True)
    w2, _ = ob.msm_g2(b2, sc[:40 * 32])
    assert ob.msm_closed_form(sc[:40 * 32], 40, 9, g2=True) == w2
    assert ob.set_threads(0) >= 1
