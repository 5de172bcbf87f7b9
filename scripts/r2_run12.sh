#!/bin/bash
# round 2, call 12: small-batch routing, wavefront skip, gather probe (DRAM bytes per miss), three-mode parity tests
python -m pytest tests -m gpu -x -q > gpurun_out/r2l_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2l_tests.log
python scripts/small_call_probe.py > gpurun_out/r2l_small.log 2>&1; echo "small rc=$?"; cat gpurun_out/r2l_small.log
python bench.py --steps 5 --warmup 3 --no-legs --cpu-seconds 3 --inexact-queries 0 > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; echo "bench rc=$?"
python - <<'P'
import json
d = [json.loads(l) for l in open("gpurun_out/r2l_bench.json") if l.startswith("{")][-1]
w = d["work"]
print("value %.1fM ms %.2f e2e %.1fM sga %.2f fast %.2f stages %s gather %s lf %s" % (d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6, w["ms_sga_kernel"], w["ms_sw_fast_kernels"], w["ms_stage"], d["roofline_sga"]["random_gather_peak_gbs"], d["lf_walk"]["frac_of_random_gather"]))
P
./scripts/probes/gather_probe > gpurun_out/r2l_gather_probe.log 2>&1; echo "probe rc=$?"; cat gpurun_out/r2l_gather_probe.log
./scripts/probes/gather_probe 32 > gpurun_out/r2l_gather_probe32.log 2>&1; tail -11 gpurun_out/r2l_gather_probe32.log
ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum --clock-control none --csv --log-file gpurun_out/r2l_gather_ncu.csv ./scripts/probes/gather_probe > /dev/null 2>&1; echo "ncu rc=$?"
python - <<'P'
import csv
rows = [r for r in csv.reader(open("gpurun_out/r2l_gather_ncu.csv")) if len(r) > 10]
h = rows[0]; ki, mi, vi = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
cur = {}
for r in rows[1:]:
    cur.setdefault(r[0], {"k": r[ki][:40]})[r[mi]] = r[vi]
for i, d in cur.items():
    print(i, d)
P
