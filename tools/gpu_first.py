"""First GPU trip: integer-pipe microbenchmark + quick MSM/NTT timings (writes JSON to gpurun_out/)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import eigen_zeth_b200 as z

ctx = z.Context(0)
out = {"int_pipe": [ctx.bench_int_pipe(4000) for _ in range(3)]}
print(out, flush=True)
stream = torch.cuda.ExternalStream(ctx.stream)

def timed(fn, reps=3):
    fn(); ctx.sync()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); fn(); e1.record(stream); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return ts

for log_n in (16, 20, 22, 24):
    n = 1 << log_n
    x = torch.randint(0, 2**31 - 1, (n, 8), dtype=torch.int32, device="cuda")
    x[:, 7] &= 0x0FFFFFFF
    for full in (0, 1):
        ctx.set("ntt_full_twiddles", full)
        ts = timed(lambda: ctx.fr_ntt_dev(x.data_ptr(), x.data_ptr(), log_n, montgomery=True))
        out["ntt_%d_full%d_ms" % (log_n, full)] = ts
        print("ntt", log_n, full, ts, flush=True)
    ctx.set("ntt_full_twiddles", 0)
for log_n in (16, 20, 22, 24):
    n = 1 << log_n
    t0 = time.time(); bases = ctx.synthetic_bases(n, 0x5EED0001); print("bases", log_n, time.time() - t0, flush=True)
    sc = torch.randint(0, 2**31 - 1, (n, 8), dtype=torch.int32, device="cuda")
    sc[:, 7] &= 0x0FFFFFFF
    res = torch.zeros(64, dtype=torch.uint8, device="cuda")
    for w in (0, 13, 16):
        ctx.set("msm_window", w)
        ts = timed(lambda: ctx.msm_dev(bases, sc.data_ptr(), n, res.data_ptr()))
        out["msm_%d_c%d_ms" % (log_n, w)] = ts
        print("msm", log_n, w, ts, flush=True)
    ctx.set("msm_window", 0)
    bases.free()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/first.json", "w"), indent=1)
