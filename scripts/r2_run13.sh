#!/bin/bash
# round 2, call 13: sga_kernel with the next read's fused entries prefetched (A/B of the modes), launch list of a 200-read call
B="python bench.py --steps 5 --warmup 3 --no-legs --no-side --no-e2e --cpu-seconds 3"
for m in 0 1 2 3; do
  BWT_SGA_PREFETCH=$m $B > gpurun_out/r2m_pf$m.json 2> gpurun_out/r2m_pf$m.err; echo "pf$m rc=$?"
done
python - <<'P'
import json
for m in range(4):
    try:
        d = [json.loads(l) for l in open("gpurun_out/r2m_pf%d.json" % m) if l.startswith("{")][-1]
        w = d["work"]
        print("prefetch", m, "value %.1fM ms %.2f sga %.2f fast %.2f parity %s" % (d["value"] / 1e6, d["ms_per_step"], w["ms_sga_kernel"], w["ms_sw_fast_kernels"], d.get("parity_sample", {}).get("mismatches")))
    except Exception as e:
        print(m, "failed", e)
P
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2m_small_launches.csv python scripts/small_call_once.py 200 > gpurun_out/r2m_small_ncu.log 2>&1; echo "small launch list rc=$?"
python - <<'P'
import csv
rows = [r for r in csv.reader(open("gpurun_out/r2m_small_launches.csv")) if len(r) > 10]
h = rows[0]; ki, vi, gi, bi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size"), h.index("Block Size")
tail = rows[-40:]
for r in tail:
    print(r[0], r[ki].split("(")[0][-44:], r[gi], r[bi], "%.1f us" % (float(r[vi].replace(",", "")) / 1e3))
P
for m in 1 2 3; do
CMD="python bench.py --steps 1 --warmup 1 --cpu-seconds 0 --no-e2e --no-side --no-legs"
BWT_SGA_PREFETCH=$m ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:sga_kernel -s 1 -c 1 --csv --log-file gpurun_out/r2m_sga_pf$m.csv $CMD > /dev/null 2>&1; echo "ncu pf$m rc=$?"; tail -3 gpurun_out/r2m_sga_pf$m.csv | cut -d, -f5,13-
done
