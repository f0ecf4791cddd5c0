"""Static SASS evidence for the hot kernels of libzkb.so (no GPU needed): instruction mix of the whole
kernel and of its main loop, stack / spill use, and an excerpt of the loop body showing the
IMAD.WIDE.U32 carry chains.  Writes profiles/<tag>_sass_summary.md and profiles/<tag>_sass_<kernel>.txt.

  python tools/sass_report.py r02
"""
import os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "eigen-zeth_b200", "libzkb.so")
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
KERNELS = [("msm_accumulate_l1_kernel<Fq>", r"msm_accumulate_l1_kernelINS_2FpINS_8FqParams"),
           ("msm_accumulate_l1_kernel<Fq2>", r"msm_accumulate_l1_kernelINS_4Fq2TILb0E"),
           ("msm_accumulate_l1_kernel<Fq2L>", r"msm_accumulate_l1_kernelINS_4Fq2LE"),
           ("ntt_pass_kernel<false>", r"ntt_pass_kernelILb0E"),
           ("ntt_pass_kernel<true>", r"ntt_pass_kernelILb1E"),
           ("sort_scatter_kernel<22>", r"sort_scatter_kernelILj22E"),
           ("sort_place_kernel", r"sort_place_kernel"),
           ("msm_reduce_level_kernel<Fq>", r"msm_reduce_level_kernelINS_2FpINS_8FqParams"),
           ("imad_chain_kernel", r"imad_chain_kernel"),
           ("montmul_chain_kernel", r"montmul_chain_kernel")]
OPS = ["IMAD.WIDE", "IMAD.HI", "IMAD", "IADD3", "LOP3", "SEL", "SHF", "SHFL", "LDG", "STG", "LDS", "STS", "LDL", "STL",
       "ISETP", "BRA", "BAR", "WARPSYNC", "MATCH", "LDC", "MOV", "PRMT", "UTMALDG", "HMMA", "UTC"]

elf = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout
names = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", names)
res_usage = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
rows = []
for label, pat in KERNELS:
    body = next((f for f in funcs if re.match(r"\S*" + pat, f)), None)
    if body is None:
        rows.append((label, None))
        continue
    mangled = body.split("\n", 1)[0].strip()
    ins = [l for l in body.split("\n") if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l)]
    addr = lambda l: int(re.match(r"\s+/\*([0-9a-f]+)\*/", l).group(1), 16)
    op = lambda l: re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", l).group(1)
    loops = []
    for l in ins:
        m = re.search(r"BRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)", l)
        if m and int(m.group(1), 16) < addr(l):
            loops.append((int(m.group(1), 16), addr(l)))
    main = max(loops, key=lambda t: t[1] - t[0]) if loops else None

    def mix(lines):
        c = {}
        for l in lines:
            o = op(l)
            key = next((k for k in OPS if o == k or o.startswith(k + ".") or (k in ("UTC",) and o.startswith(k))), "other")
            if o.startswith("IMAD.WIDE"):
                key = "IMAD.WIDE"
            elif o.startswith("IMAD.HI"):
                key = "IMAD.HI"
            c[key] = c.get(key, 0) + 1
        return c

    whole = mix(ins)
    loop_ins = [l for l in ins if main and main[0] <= addr(l) <= main[1]]
    lm = mix(loop_ins) if loop_ins else {}
    usage = re.search(re.escape(mangled) + r":\s*\n\s*(REG:\d+[^\n]*)", res_usage)
    rows.append((label, {"mangled": mangled, "n": len(ins), "whole": whole, "loop_n": len(loop_ins), "loop": lm,
                         "usage": usage.group(1) if usage else "?"}))
    fn = "%s_sass_%s.txt" % (tag, re.sub(r"[^A-Za-z0-9]+", "_", label).strip("_"))
    with open(os.path.join(ROOT, "profiles", fn), "w") as f:
        f.write("# %s (%s)\n# cuobjdump -sass eigen-zeth_b200/libzkb.so ; %d instructions, main loop %s (%d instructions)\n"
                % (label, mangled, len(ins), "0x%x..0x%x" % main if main else "-", len(loop_ins)))
        f.write("# resources: %s\n# excerpt: first 160 instructions of the main loop (carry chains = IMAD.WIDE.U32 / IMAD.WIDE.U32.X runs)\n" % (usage.group(1) if usage else "?"))
        for l in (loop_ins or ins)[:160]:
            f.write(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", l).rstrip() + "\n")

with open(os.path.join(ROOT, "profiles", "%s_sass_summary.md" % tag), "w") as f:
    f.write("# Static SASS evidence (%s) — `python tools/sass_report.py %s` on the built `libzkb.so`\n\n" % (tag, tag))
    f.write("cubins in the library: %s\n\n" % ", ".join(sorted(set(re.findall(r"sm_\d+a?", elf)))))
    f.write("| kernel | instructions | IMAD.WIDE | other IMAD | IADD3 | LOP3+SEL+SHF | SHFL | LDL/STL (spill) | main loop: instr / IMAD.WIDE | resources |\n|---|---|---|---|---|---|---|---|---|---|\n")
    for label, r in rows:
        if r is None:
            f.write("| `%s` | not found | | | | | | | | |\n" % label)
            continue
        w, l = r["whole"], r["loop"]
        f.write("| `%s` | %d | %d | %d | %d | %d | %d | %d | %d / %d | %s |\n" % (
            label, r["n"], w.get("IMAD.WIDE", 0), w.get("IMAD", 0) + w.get("IMAD.HI", 0), w.get("IADD3", 0),
            w.get("LOP3", 0) + w.get("SEL", 0) + w.get("SHF", 0), w.get("SHFL", 0), w.get("LDL", 0) + w.get("STL", 0),
            r["loop_n"], l.get("IMAD.WIDE", 0), r["usage"]))
    f.write("\nNo `HMMA` / `UTC*MMA` / `UTMALDG` anywhere: this is big-integer work on the FMA-heavy integer pipe, "
            "tensor cores and TMA are not used (BASELINE.json north_star).  Loop excerpts: `profiles/%s_sass_*.txt`.\n" % tag)
print(open(os.path.join(ROOT, "profiles", "%s_sass_summary.md" % tag)).read())
