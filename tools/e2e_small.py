"""small end-to-end pass over every kernel family (quick health check; written for a
compute-sanitizer memcheck run, but that tool is closed on this pool): NTT (1, 2 and 3 passes, all variants), batched NTT + permute, G1/G2 MSM (classic,
precomputed, chunked host path), Groth16 prove 2^8 + pairing check."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import eigen_zeth_b200 as z
import oracle_b as ob
ctx = z.Context(0)
for log_n in (5, 10, 17):
    v = ob.random_scalars(1 << log_n, log_n)
    for inv in (False, True):
        for coset in (False, True):
            out = ctx.fr_ntt(v, inverse=inv, coset=coset)
            if log_n <= 10:
                assert out == ob.fr_ntt(v, inv, coset)
n = 777
b1 = ob.synthetic_bases(n, 1); b2 = ob.synthetic_bases(200, 2, g2=True); sc = ob.random_scalars(n, 3)
B1 = ctx.upload_bases(b1); B2 = ctx.upload_bases(b2, g2=True)
want1, _ = ob.msm_g1(b1, sc); want2, _ = ob.msm_g2(b2, sc[:200 * 32])
assert ctx.msm(B1, sc) == want1 and ctx.msm(B2, sc[:200 * 32]) == want2
ctx.set("msm_host_chunks", 3)
assert ctx.msm(B1, sc) == want1
ctx.set("msm_host_chunks", 0)
B1.precompute(0); B2.precompute(0)
assert ctx.msm(B1, sc) == want1 and ctx.msm(B2, sc[:200 * 32]) == want2
pk = ctx.pk_synthetic(8, 5)
w = z.synthetic_witness(8, (3).to_bytes(32, "little"), (4).to_bytes(32, "little"))
proof, _, _, _ = ctx.groth16_prove(pk, w, (1).to_bytes(32, "little"), (2).to_bytes(32, "little"))
assert z.groth16_verify(pk.vk, proof, w[32:64])
print("E2E-SMALL OK")
