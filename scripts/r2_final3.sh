#!/bin/bash
# round 2, final call of the band / pipeline work: full GPU suite, default bench line, reference arm, launch list and
# ncu --set full captures of the step's kernels (tag r2fin3)
T=r2fin3
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${T}_tests.log
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
python bench.py --edits sub --steps 5 --warmup 3 --no-legs --cpu-seconds 3 > gpurun_out/${T}_bench_sub.json 2> gpurun_out/${T}_bench_sub.err; echo "bench sub rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; echo "bench ref rc=$?"
CMD="python bench.py --steps 1 --warmup 1 --cpu-seconds 0 --no-e2e --no-side --no-legs"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches.csv $CMD > gpurun_out/${T}_ncu_launch.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'sga_kernel|sw_classify_kernel|sw_wfa_kernel|sw_band_kernel|sw_fast_kernel|emit_kernel' -s 6 -c 6 -o gpurun_out/prof_${T} -f $CMD > gpurun_out/${T}_ncu.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/${T}_ncu.log
python - <<'P'
import json
for f in ["r2fin3_bench", "r2fin3_bench_sub", "r2fin3_bench_ref"]:
    try:
        d = [json.loads(l) for l in open("gpurun_out/%s.json" % f) if l.startswith("{")][-1]
        w = d.get("work") or {}
        print(f, "value %.1fM ms %.2f e2e %.1fM" % (d["value"] / 1e6, d["ms_per_step"], (d.get("e2e") or {}).get("value", 0) / 1e6), "sga", w.get("ms_sga_kernel"), "fast", w.get("ms_sw_fast_kernels"), "band", w.get("sw_band_pairs_frac"), "parity", (d.get("parity_sample") or {}).get("mismatches"))
    except Exception as e:
        print(f, "failed", e)
P
