"""profiles/traffic.json from an `ncu --set full` raw CSV (`ncu -i X.ncu-rep --page raw --csv`):
dram__bytes_read.sum / dram__bytes_write.sum per captured launch of the dominant kernel, so that
bench.py's roofline.traffic is read from a committed capture rather than typed in.

  python tools/update_traffic.py profiles/r02_ncu_full_msm_accumulate_l1_raw.csv 24 22 <commit>
"""
import csv, json, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path, log_n, window, commit = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
rows = list(csv.reader(open(path)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
caps = []
for d in data:
    name = d[idx["Kernel Name"]]
    if "msm_accumulate_l1_kernel" not in name:
        continue
    kern = "msm_accumulate_l1_kernel<Fq2>" if "Fq2" in name else "msm_accumulate_l1_kernel<Fq>"
    rd = float(d[idx["dram__bytes_read.sum"]]) * scale[units[idx["dram__bytes_read.sum"]]]
    wr = float(d[idx["dram__bytes_write.sum"]]) * scale[units[idx["dram__bytes_write.sum"]]]
    caps.append({"kernel": kern, "log_n": log_n, "window_bits": window, "dram_bytes_read": int(rd),
                 "dram_bytes_write": int(wr), "gpu_time_ms": float(d[idx["gpu__time_duration.sum"]]),
                 "fmaheavy_pct": float(d[idx["sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"]]),
                 "source": "%s (ncu --set full, commit %s)" % (os.path.relpath(path, ROOT), commit)})
if not caps:
    sys.exit("no msm_accumulate_l1_kernel launch in " + path)
out = os.path.join(ROOT, "profiles", "traffic.json")
doc = json.load(open(out))
keep = [c for c in doc["captures"] if not any(c["kernel"] == n["kernel"] and c["log_n"] == n["log_n"] and
                                              c["window_bits"] == n["window_bits"] for n in caps)]
doc["captures"] = keep + caps[-1:]
json.dump(doc, open(out, "w"), indent=1)
print(json.dumps(caps[-1], indent=1))
