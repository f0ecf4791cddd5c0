"""K5: the library's own bucket sort (digit extraction fused, two-level MSD) against digits kernel +
cub::DeviceRadixSort, same scalars, stage times from zkb_msm_last_timings (CUDA events)."""
import json, os, statistics, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import eigen_zeth_b200 as z
import oracle_b as ob

sizes = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "19,22,24").split(",")]
windows = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "0").split(",")]   # 0 = the library's choice
ctx = z.Context(0)
out = []
for log_n, window in [(a, b) for a in sizes for b in windows]:
    n = 1 << log_n
    bases = ctx.synthetic_bases(n, 0x5EED0001)
    bases.precompute(window)
    res = torch.zeros(64, dtype=torch.uint8, device="cuda")
    uni = ob.random_scalars(n, 0x5EED0000 + log_n)
    inputs = {"uniform": uni, "zero_one": b"".join((i & 1).to_bytes(32, "little") for i in range(1 << 16)) * (n >> 16),
              "all_equal": uni[:32] * n}
    for name, sc in inputs.items():
        dev = torch.frombuffer(bytearray(sc), dtype=torch.uint8).cuda()
        want = ob.msm_closed_form(sc, n, 0x5EED0001)
        row = {"log_n": log_n, "scalars": name, "window_bits": bases.window}
        for mode, label in ((1, "cub"), (0, "bucket_sort")):
            ctx.set("msm_sort", mode)
            ctx.set("msm_profile", 1)
            st = {}
            for i in range(5):
                ctx.msm_dev(bases, dev.data_ptr(), n, res.data_ptr())
                t = ctx.msm_last_timings()
                if i:
                    for k, v in t.items():
                        st.setdefault(k, []).append(v)
            ctx.sync()
            row[label] = {k: round(statistics.median(v), 3) for k, v in st.items()}
            row[label]["digits_plus_sort_ms"] = round(row[label]["digits"] + row[label]["sort"], 3)
            row[label]["msm_ms"] = round(sum(statistics.median(v) for v in st.values()), 3)
            row[label]["correct"] = bytes(res.cpu().numpy()) == want
        ctx.set("msm_profile", 0)
        ctx.set("msm_sort", 0)
        out.append(row)
        print(json.dumps(row), flush=True)
    bases.free()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "sort_time.json"), "w"), indent=1)
