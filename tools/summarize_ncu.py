#!/usr/bin/env python
"""Turn the ncu outputs that gpurun brought back (gpurun_out/) into the tracked summaries under
profiles/: a per-kernel launch table (share of the step) and the key counters of each full capture.
    python tools/summarize_ncu.py r1 dna_100x100k
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
wl = sys.argv[2] if len(sys.argv) > 2 else "dna_100x100k"
G = os.path.join(ROOT, "gpurun_out")
PDIR = os.path.join(ROOT, "profiles")
os.makedirs(PDIR, exist_ok=True)
out = [f"# ncu summary {tag} / {wl}", "",
       "Source: `tools/profile.sh` on one B200 (`bench.py --steps 2 --warmup 3`), read back with "
       "`ncu -i ... --page raw --csv`.  Launch times under ncu are cold-cache and serialised: use the SHARES."]

lst = os.path.join(G, f"launches_{tag}_{wl}.csv")
if os.path.exists(lst):
    rows = [r for r in csv.reader(open(lst)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    d = collections.OrderedDict()
    for r in rows[1:]:
        try:
            d.setdefault(re.sub(r"\(.*", "", r[ki]), []).append(float(r[vi].replace(",", "")))
        except ValueError:
            pass
    tot = sum(sum(v) for v in d.values())
    out += ["", "## launch list (gpu__time_duration.sum)", "", "| kernel | launches | total us | avg us | share |", "|---|---|---|---|---|"]
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        out.append(f"| `{k}` | {len(v)} | {sum(v)/1e3:.1f} | {sum(v)/len(v)/1e3:.2f} | {100*sum(v)/tot:.1f}% |")
    with open(os.path.join(PDIR, f"launches_{tag}_{wl}.csv"), "w") as f:
        f.write("kernel,launches,total_us,avg_us,share\n")
        for k, v in d.items():
            f.write(f"\"{k}\",{len(v)},{sum(v)/1e3:.3f},{sum(v)/len(v)/1e3:.3f},{sum(v)/tot:.4f}\n")

PAT = re.compile(r"^(gpu__time_duration.sum|dram__bytes_read.sum|dram__bytes_write.sum|dram__throughput.avg.pct_of_peak_sustained_elapsed|"
                 r"gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed|sm__warps_active.avg.pct_of_peak_sustained_active|launch__registers_per_thread|"
                 r"launch__grid_size|launch__block_size|launch__occupancy_limit_registers|lts__t_sector_hit_rate.pct|l1tex__t_sector_hit_rate.pct|"
                 r"sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active|sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active|"
                 r"sm__inst_executed_pipe_tensor.*pct.*|smsp__inst_executed.sum|smsp__issue_active.avg.pct_of_peak_sustained_active|"
                 r"sm__throughput.avg.pct_of_peak_sustained_elapsed|l1tex__data_pipe_lsu_wavefronts_mem_shared.sum|sm__cycles_elapsed.max)$")
traffic = {}
for kern in ("k1", "k2", "k0"):
    rep = os.path.join(G, f"prof_{tag}_{wl}_{kern}.ncu-rep")
    if not os.path.exists(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        continue
    hdr, units, vals = rows[0], rows[1], rows[2]
    name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else kern
    out += ["", f"## full capture: {kern} = `{name[:110]}`", "", "| metric | value | unit |", "|---|---|---|"]
    rd = wr = None
    for i, h in enumerate(hdr):
        if PAT.match(h):
            out.append(f"| {h} | {vals[i]} | {units[i]} |")
        if h == "dram__bytes_read.sum":
            rd = (float(vals[i]), units[i])
        if h == "dram__bytes_write.sum":
            wr = (float(vals[i]), units[i])
    if rd and wr:
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        traffic[kern] = rd[0] * mult.get(rd[1], 1) + wr[0] * mult.get(wr[1], 1)
        out.append(f"| **traffic = dram read + write per launch** | {traffic[kern]/1e6:.1f} | MB |")
    # stall reasons from the source page
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(src.splitlines()))
    if len(srows) > 3:
        h2 = srows[1]
        ci = {h: i for i, h in enumerate(h2)}
        st = [h for h in h2 if h.startswith("stall_") and "Not Issued" not in h]
        agg = {h: sum(int(r[ci[h]] or 0) for r in srows[2:] if len(r) == len(h2)) for h in st}
        tot = sum(agg.values()) or 1
        top = sorted(agg.items(), key=lambda kv: -kv[1])[:6]
        out.append("")
        out.append("warp-stall samples: " + ", ".join(f"{k[6:]} {100*v/tot:.0f}%" for k, v in top))
        ops = collections.Counter()
        for r in srows[2:]:
            if len(r) != len(h2):
                continue
            m = r[ci["Source"]].strip().split()
            if not m:
                continue
            nm = m[1] if m[0].startswith("@") and len(m) > 1 else m[0]
            ops[nm.split(".")[0]] += int(r[ci["Instructions Executed"]] or 0)
        out.append("")
        out.append("executed SASS by opcode (warp-level): " + ", ".join(f"{k} {v}" for k, v in ops.most_common(14)))

if "k1" in traffic:
    tp = os.path.join(PDIR, "traffic.json")
    cur = json.load(open(tp)) if os.path.exists(tp) else {}
    cur[wl] = traffic["k1"]
    json.dump(cur, open(tp, "w"), indent=1)
open(os.path.join(PDIR, f"ncu_{tag}_{wl}.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
