"""time the precomputed-bases MSM for a grid of (size, window width) to calibrate msm_auto_window."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import eigen_zeth_b200 as z
ctx = z.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream)
out = {}
for log_n in (14, 16, 18, 19, 20, 21, 22, 23, 24):
    n = 1 << log_n
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import uniform_fr_scalars
    sc = torch.from_numpy(uniform_fr_scalars(n, log_n).view("int64")).cuda()
    res = torch.zeros(64, dtype=torch.uint8, device="cuda")
    for c in range(max(8, log_n - 7), min(23, log_n + 1) + 1):
        bases = ctx.synthetic_bases(n, 1)
        try:
            bases.precompute(c)
        except z.ZkbError as e:
            bases.free(); continue
        for _ in range(2):
            ctx.msm_dev(bases, sc.data_ptr(), n, res.data_ptr())
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); ctx.msm_dev(bases, sc.data_ptr(), n, res.data_ptr()); e1.record(stream); e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        out["%d,%d" % (log_n, c)] = sorted(ts)[2]
        bases.free()
    best = min((v, k) for k, v in out.items() if k.startswith("%d," % log_n))
    print(log_n, {k.split(",")[1]: round(v, 3) for k, v in out.items() if k.startswith("%d," % log_n)}, "best", best, flush=True)
json.dump(out, open("gpurun_out/tune_window.json", "w"))
