#!/usr/bin/env python
"""Turn gpurun_out/<launches>.csv and gpurun_out/<prof>.ncu-rep into the tracked summaries under profiles/.

  python scripts/ncu_summary.py r1a gpurun_out/launches.csv gpurun_out/prof_r1.ncu-rep "<command line>"
"""
import collections
import csv
import json
import os
import subprocess
import sys

tag, launches, rep, cmd = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

rows = [r for r in csv.reader(open(launches)) if len(r) > 10] if launches != "-" else [["ID", "Kernel Name", "Metric Value", "Grid Size", "Block Size"]]
hdr = rows[0]
ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
with open(os.path.join(out_dir, "%s_launches.csv" % tag) if launches != "-" else os.devnull, "w") as f:
    f.write("# %s\n# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare shares)\n" % cmd)
    f.write("id,kernel,grid,block,time_us\n")
    for r in rows[1:]:
        f.write("%s,%s,%s,%s,%.1f\n" % (r[0], r[ki].split("(")[0].replace("<unnamed>::", "").replace("void ", "")[-48:],
                                        r[gi].replace(",", " "), r[bi].replace(",", " "), float(r[vi].replace(",", "")) / 1e3))
agg = collections.OrderedDict()
for r in rows[1:]:
    name = r[ki].split("(")[0].replace("<unnamed>::", "").replace("void ", "")[-48:]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi].replace(",", "")) / 1e6

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, units = rr[0], rr[1]
idx = {n: i for i, n in enumerate(h)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
kern = []
for r in rr[2:]:
    d = {"kernel": r[idx["Kernel Name"]].split("(")[0].replace("<unnamed>::", "").replace("void ", "")}
    for w in want:
        if w in idx:
            try:
                d[w] = float(r[idx[w]].replace(",", ""))
            except ValueError:
                d[w] = r[idx[w]]
            d[w + "@unit"] = units[idx[w]]
    kern.append(d)

with open(os.path.join(out_dir, "%s_summary.md" % tag), "w") as f:
    f.write("# ncu summary %s\n\ncommand: `%s`\n\n## launch list (per kernel, all launches of the command)\n\n| kernel | launches | total ms |\n|---|---|---|\n" % (tag, cmd))
    for k, v in agg.items():
        f.write("| %s | %d | %.3f |\n" % (k, v[0], v[1]))
    f.write("\n## `ncu --set full` captures\n")
    for d in kern:
        f.write("\n### %s\n\n| metric | value | unit |\n|---|---|---|\n" % d["kernel"])
        for w in want:
            if w in d:
                f.write("| %s | %s | %s |\n" % (w, d[w], d[w + "@unit"]))
        if "dram__bytes_read.sum" in d:
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
            tot = d["dram__bytes_read.sum"] * scale[d["dram__bytes_read.sum@unit"]] + d["dram__bytes_write.sum"] * scale[d["dram__bytes_write.sum@unit"]]
            tms = d["gpu__time_duration.sum"] * {"ms": 1, "us": 1e-3, "s": 1e3, "ns": 1e-6}[d["gpu__time_duration.sum@unit"]]
            f.write("| DRAM traffic per launch | %.4g | byte |\n| DRAM GB/s under ncu | %.1f | GB/s |\n" % (tot, tot / tms / 1e6))
            d["dram_bytes_per_launch"] = tot
# per-launch DRAM traffic of the big launches, for bench.py's roofline.traffic (largest launch of each kernel)
traffic = {}
for d in kern:
    name = d["kernel"].split("<")[0]
    if "dram_bytes_per_launch" in d and d["dram_bytes_per_launch"] > traffic.get(name, {}).get("dram_bytes_per_launch", 0):
        traffic[name] = {"dram_bytes_per_launch": d["dram_bytes_per_launch"], "ms_under_ncu": d["gpu__time_duration.sum"],
                         "source": "profiles/%s_summary.md (ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum; `%s`)" % (tag, cmd)}
if len(sys.argv) > 5 and sys.argv[5] == "--traffic":
    with open(os.path.join(out_dir, "ncu_traffic.json"), "w") as f:
        json.dump(traffic, f, indent=1)
print(json.dumps([{k: v for k, v in d.items() if "@" not in k} for d in kern], indent=1))
