"""G2 MSM variants behind the "msm_g2_pair" switch: bit 0 = bucket WALK with one Fq2 point operation per
lane pair (Fq2L, 16 warps / SM) instead of one thread per point (255 registers, 8 warps / SM); bit 1 =
the latency-bound kernels around the walk on lane pairs instead of out-of-line Fq2 products.
Same bases, same scalars, same result."""
import json, os, statistics, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import eigen_zeth_b200 as z

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 22
n = 1 << log_n
ctx = z.Context(0)
bases = ctx.synthetic_bases(n, 0x5EED0002, g2=True)
bases.precompute(0)
g = torch.Generator().manual_seed(1)
sc = torch.randint(0, 2**62, (n, 4), dtype=torch.int64, generator=g)
sc[:, 3] &= (1 << 60) - 1
sc = sc.cuda()
out = torch.zeros(128, dtype=torch.uint8, device="cuda")
stream = torch.cuda.ExternalStream(ctx.stream)
res = {}
pipe = ctx.bench_int_pipe(3000)
ctx.set("msm_profile", 1)
ctx.set("msm_g2_pair", 2)
MODES = {0: "walk_single_tails_calls", 2: "walk_single_tails_pair", 1: "walk_pair_tails_calls", 3: "walk_pair_tails_pair"}
for pair in (0, 2, 1, 3, 0, 2):
    ctx.set("msm_g2_pair", pair)
    for _ in range(2):
        ctx.msm_dev(bases, sc.data_ptr(), n, out.data_ptr())
    ctx.sync()
    ts, l1 = [], []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ctx.msm_dev(bases, sc.data_ptr(), n, out.data_ptr())
        e1.record(stream)
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
        l1.append(ctx.msm_last_timings()["accumulate_l1"])
    key = MODES[pair]
    res.setdefault(key, {"msm_ms": [], "walk_ms": [], "result": None})
    res[key]["msm_ms"].append(statistics.median(ts))
    res[key]["walk_ms"].append(statistics.median(l1))
    res[key]["result"] = bytes(out.cpu().numpy()).hex()
    res[key]["stage_ms"] = {k: round(v, 3) for k, v in ctx.msm_last_timings().items()}
c = bases.window
nwin = (255 + c - 1) // c
imads = n * nwin * (8 * 400 + 2 * 272)          # 8 Fq2 products + 2 Fq2 squarings per mixed addition
for k, v in list(res.items()):
    v["walk_frac_int_pipe"] = imads / (min(v["walk_ms"]) * 1e-3) / pipe["imad_wide_per_s"]
res["same_result"] = len({v["result"] for v in res.values()}) == 1
res["window_bits"], res["windows"], res["log_n"], res["imad_wide_per_s"] = c, nwin, log_n, pipe["imad_wide_per_s"]
for k in MODES.values():
    res[k]["result"] = res[k]["result"][:16] + "..."
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "g2_pair_time_2p%d.json" % log_n), "w"), indent=1)
print(json.dumps(res))
