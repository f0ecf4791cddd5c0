"""BASELINE.json configs[1]: G1 MSM + Fr NTT size sweep 2^16..2^26 on one B200 vs the CPU oracle on
the host cores.  Bit-exactness at EVERY size: the MSM against the closed form (sum s_i k_i mod r) * G
(oracle B: one dot product + one scalar multiplication) and, up to 2^SWEEP_CPU_MAX, against oracle B's
full Pippenger as well (that one is also the timed CPU column); the NTT against oracle B's transform.
Writes gpurun_out/sweep.json/.md"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch  # noqa: E402
import eigen_zeth_b200 as z  # noqa: E402
import oracle_b as ob  # noqa: E402

CPU_MAX = int(os.environ.get("SWEEP_CPU_MAX", "22"))
ob.set_threads(0)
sizes = [int(x) for x in os.environ.get("SWEEP_SIZES", "16,18,20,22,24,26").split(",")]
ctx = z.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream)
pipe = ctx.bench_int_pipe(3000)
rows = []


def timed(fn, reps):
    fn()
    ctx.sync()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


for log_n in sizes:
    n = 1 << log_n
    seed = 0x5EED0000 + log_n
    row = {"log_n": log_n}
    reps = 10 if log_n <= 24 else 3
    # ---- inputs (same generators as the CPU side)
    scal = ob.random_scalars(n, seed)
    sc_dev = torch.frombuffer(bytearray(scal), dtype=torch.uint8).cuda()
    # ---- NTT
    x = sc_dev.clone()
    ctx.fr_convert_dev(x.data_ptr(), n, True)
    row["ntt_ms"] = timed(lambda: ctx.fr_ntt_dev(x.data_ptr(), x.data_ptr(), log_n, montgomery=True), reps)
    row["ntt_gelem_s"] = n / row["ntt_ms"] / 1e6
    row["ntt_frac_int_pipe"] = 136 * (n // 2) * log_n / (row["ntt_ms"] * 1e-3) / pipe["imad_wide_per_s"]
    row["ntt_frac_hbm"] = 128 * n / (row["ntt_ms"] * 1e-3) / 1e9 / 6533.2
    del x
    # ---- MSM (precomputed window multiples, device-resident scalars)
    bases = ctx.synthetic_bases(n, seed + 1)
    res = torch.zeros(64, dtype=torch.uint8, device="cuda")
    ctx.msm_dev(bases, sc_dev.data_ptr(), n, res.data_ptr())
    ctx.sync()
    plain_result = bytes(res.cpu().numpy())
    c = bases.precompute(0)
    nwin = (255 + c - 1) // c
    row["msm_window"], row["msm_windows"] = c, nwin
    row["msm_ms"] = timed(lambda: ctx.msm_dev(bases, sc_dev.data_ptr(), n, res.data_ptr()), reps)
    row["msm_mpts_s"] = n / row["msm_ms"] / 1e3
    row["msm_frac_int_pipe"] = n * nwin * 1360 / (row["msm_ms"] * 1e-3) / pipe["imad_wide_per_s"]
    ctx.sync()
    gpu_result = bytes(res.cpu().numpy())
    row["msm_plain_equals_precomputed"] = gpu_result == plain_result
    row["msm_bit_exact_closed_form"] = gpu_result == ob.msm_closed_form(scal, n, seed + 1)
    ctx.set("msm_profile", 1)
    ctx.msm_dev(bases, sc_dev.data_ptr(), n, res.data_ptr())
    row["msm_stage_ms"] = {k: round(v, 3) for k, v in ctx.msm_last_timings().items()}
    ctx.set("msm_profile", 0)
    # ---- CPU
    if log_n <= CPU_MAX:
        t0 = time.time()
        cb = ob.synthetic_bases(n, seed + 1)
        row["cpu_bases_s"] = time.time() - t0
        cpu_res, t = ob.msm_g1(cb, scal)
        row["cpu_msm_s"], row["cpu_msm_mpts_s"] = t, n / t / 1e6
        row["msm_bit_exact"] = cpu_res == gpu_result
        row["cpu_threads"] = ob.threads()
    cpu_ntt, t = ob.fr_ntt_timed(scal)
    row["cpu_ntt_s"], row["cpu_ntt_gelem_s"] = t, n / t / 1e9
    row["ntt_bit_exact"] = cpu_ntt == ctx.fr_ntt(scal)
    del cpu_ntt
    bases.free()
    del sc_dev
    torch.cuda.empty_cache()
    rows.append(row)
    print(row, flush=True)

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump({"int_pipe": pipe, "rows": rows}, open(os.path.join(ROOT, "gpurun_out", "sweep.json"), "w"), indent=1)
with open(os.path.join(ROOT, "gpurun_out", "sweep.md"), "w") as f:
    f.write("| size | MSM ms | Mpts/s | c | frac IMAD | NTT ms | Gelem/s | frac IMAD | frac HBM | CPU MSM s (thr) | CPU NTT s | MSM bit-exact | NTT bit-exact |\n|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
    for r in rows:
        f.write("| 2^%d | %.3f | %.1f | %d | %.2f | %.3f | %.2f | %.2f | %.3f | %s | %s | %s | %s |\n" % (
            r["log_n"], r["msm_ms"], r["msm_mpts_s"], r["msm_window"], r["msm_frac_int_pipe"], r["ntt_ms"], r["ntt_gelem_s"],
            r["ntt_frac_int_pipe"], r["ntt_frac_hbm"],
            ("%.2f (%d)" % (r["cpu_msm_s"], r["cpu_threads"])) if "cpu_msm_s" in r else "-",
            ("%.3f" % r["cpu_ntt_s"]) if "cpu_ntt_s" in r else "-",
            "%s (closed form)%s" % (r["msm_bit_exact_closed_form"], (", %s (full CPU MSM)" % r["msm_bit_exact"]) if "msm_bit_exact" in r else ""),
            r.get("ntt_bit_exact", "-")))
