#!/bin/bash
# round 2, final call: full GPU suite, default bench line, launch list and ncu --set full captures of the step's kernels
python -m pytest tests -m gpu -x -q > gpurun_out/r2fin_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2fin_tests.log
python bench.py > gpurun_out/r2fin_bench.json 2> gpurun_out/r2fin_bench.err; echo "bench rc=$?"
python bench.py --edits sub --steps 5 --warmup 3 --no-legs --cpu-seconds 3 > gpurun_out/r2fin_bench_sub.json 2> gpurun_out/r2fin_bench_sub.err; echo "bench sub rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2fin_bench_ref.json 2> gpurun_out/r2fin_bench_ref.err; echo "bench ref rc=$?"
CMD="python bench.py --steps 1 --warmup 1 --cpu-seconds 0 --no-e2e --no-side --no-legs"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2fin_launches.csv $CMD > gpurun_out/r2fin_ncu_launch.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'sga_kernel|sw_classify_kernel|sw_wfa_kernel|sw_fast_kernel|emit16_kernel|emit_kernel' -s 5 -c 5 -o gpurun_out/prof_r2fin -f $CMD > gpurun_out/r2fin_ncu.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/r2fin_ncu.log
python - <<'P'
import json
for f in ["r2fin_bench", "r2fin_bench_sub", "r2fin_bench_ref"]:
    try:
        d = [json.loads(l) for l in open("gpurun_out/%s.json" % f) if l.startswith("{")][-1]
        w = d.get("work") or {}
        print(f, "value %.1fM ms %.2f e2e %.1fM" % (d["value"] / 1e6, d["ms_per_step"], (d.get("e2e") or {}).get("value", 0) / 1e6), "sga", w.get("ms_sga_kernel"), "fast", w.get("ms_sw_fast_kernels"), "parity", (d.get("parity_sample") or {}).get("mismatches"))
    except Exception as e:
        print(f, "failed", e)
P
