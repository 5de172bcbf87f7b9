#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples of one kernel from an .ncu-rep (captured with --import-source on).
  python scripts/ncu_lines.py gpurun_out/prof.ncu-rep regex:sga_kernel [top N]"""
import csv, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None
hdr = None
agg = collections.OrderedDict()
tot_inst = tot_samp = 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r; ii = hdr.index("Instructions Executed"); si = hdr.index("# Samples"); ti = hdr.index("Thread Instructions Executed"); continue
    if hdr is None or r[0] == "":
        continue          # SASS rows belong to the line above (already aggregated in the line row)
    try:
        inst = int(r[ii]); samp = int(r[si]); tinst = int(r[ti])
    except ValueError:
        continue
    key = (cur_file, int(r[0]))
    a = agg.setdefault(key, [0, 0, 0, r[1].strip()[:110]])
    a[0] += inst; a[1] += samp; a[2] += tinst
    tot_inst += inst; tot_samp += samp
print("total warp instructions %d, samples %d" % (tot_inst, tot_samp))
print("%-14s %5s %8s %6s %6s %5s  %s" % ("file", "line", "Minst", "inst%", "samp%", "thr", "source"))
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-14s %5d %8.1f %6.2f %6.2f %5.1f  %s" % (f, ln, a[0] / 1e6, 100.0 * a[0] / max(1, tot_inst), 100.0 * a[1] / max(1, tot_samp),
                                                 a[2] / max(1, a[0]), a[3]))
