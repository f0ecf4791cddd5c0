"""Window width of proving-key queries, tuned on whole proofs (one sort per witness, batched and
concurrent tails change the optimum of a lone MSM): times zkb_groth16_prove for a full key and
zkb_groth16_prove_partial for the rank-0 shard of an 8-way split, per candidate width."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import eigen_zeth_b200 as z

ctx = z.Context(0)
r, s = (0x1234567 << 200 | 5).to_bytes(32, "little"), (0x7654321 << 199 | 9).to_bytes(32, "little")
out = {}
for log_n, world, cands in ((22, 1, (19, 20, 22)), (22, 8, (15, 16, 17, 19, 20)), (24, 8, (17, 19, 20, 22)), (20, 1, (17, 19, 20))):
    w = z.synthetic_witness(log_n, (5).to_bytes(32, "little"), (7).to_bytes(32, "little"))
    wp = z.lib.host_alloc(len(w))
    z.lib.host_write(wp, w)
    for c in cands:
        ctx.set("pk_window", c)
        try:
            pk = ctx.pk_synthetic(log_n, 4) if world == 1 else ctx.pk_synthetic_shard(log_n, 4, 0, world)
        except z.ZkbError as e:
            print(log_n, world, c, "skip:", e, flush=True)
            continue
        best = None
        for i in range(4):
            if world == 1:
                _, _, _, t = ctx.groth16_prove(pk, wp, r, s)
            else:
                _, t = ctx.groth16_prove_partial(pk, wp)
            msm = sum(v for k, v in t.items() if k.startswith("ms_msm"))
            if i and (best is None or msm < best[0]):
                best = (msm, t["ms_total"])
        out["%d/%d/%d" % (log_n, world, c)] = {"msm_ms": round(best[0], 3), "total_ms": round(best[1], 3)}
        print(log_n, "world", world, "c", c, out["%d/%d/%d" % (log_n, world, c)], flush=True)
        pk.free()
    z.lib.host_free(wp)
ctx.set("pk_window", 0)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "tune_prove_window.json"), "w"), indent=1)
