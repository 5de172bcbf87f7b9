#!/bin/bash
# round 2, call 11: DP list ordered by shape (A/B), repaired gather peak, reference-golden GPU test, launch list
python -m pytest tests -m gpu -x -q > gpurun_out/r2k_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2k_tests.log
B="python bench.py --steps 5 --warmup 3 --no-legs --cpu-seconds 3 --inexact-queries 0"
BWT_SW_SORT=-1 $B --no-side > gpurun_out/r2k_nosort.json 2> gpurun_out/r2k_nosort.err; echo "nosort rc=$?"
$B > gpurun_out/r2k_sort.json 2> gpurun_out/r2k_sort.err; echo "sort rc=$?"
python bench.py --config 4 --c4-reads 20000000 --steps 2 --warmup 1 --no-legs --no-side --cpu-seconds 0 > gpurun_out/r2k_c4.json 2> gpurun_out/r2k_c4.err; echo "c4 rc=$?"
BWT_SW_SORT=-1 python bench.py --config 4 --c4-reads 20000000 --steps 2 --warmup 1 --no-legs --no-side --cpu-seconds 0 > gpurun_out/r2k_c4_nosort.json 2> gpurun_out/r2k_c4_nosort.err; echo "c4 nosort rc=$?"
python - <<'P'
import json
for f in ["r2k_nosort", "r2k_sort", "r2k_c4", "r2k_c4_nosort"]:
    try:
        d = [json.loads(l) for l in open("gpurun_out/%s.json" % f) if l.startswith("{")][-1]
        w = d.get("work", {})
        print(f, "value %.1fM ms %.2f e2e %.1fM sga %.2f fast %.2f sw %.2f" % (d["value"] / 1e6, d["ms_per_step"], d.get("e2e", {}).get("value", 0) / 1e6,
              w.get("ms_sga_kernel", 0), w.get("ms_sw_fast_kernels", 0), w.get("ms_stage", {}).get("sw", 0)),
              "gather", d.get("roofline_sga", {}).get("random_gather_peak_gbs"), "lf", d.get("lf_walk", {}).get("frac_of_random_gather"))
    except Exception as e:
        print(f, "failed", e)
P
CMD="python bench.py --steps 1 --warmup 1 --cpu-seconds 0 --no-e2e --no-side --no-legs"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2k_launches.csv $CMD > gpurun_out/r2k_ncu_launch.log 2>&1; echo "launch list rc=$?"
