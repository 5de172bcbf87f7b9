#!/bin/bash
# round 2, call 7: block-aggregated classification atomics + sharded CIGAR arena; full default bench with legs
python -m pytest tests -m gpu -x -q > gpurun_out/r2g_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2g_tests.log
SECONDS=0
python bench.py --steps 5 --warmup 3 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench rc=$? in $SECONDS s"
tail -5 gpurun_out/r2g_bench.err | cut -c1-600
python bench.py --edits sub --steps 5 --warmup 3 --no-legs --cpu-seconds 3 > gpurun_out/r2g_bench_sub.json 2> gpurun_out/r2g_bench_sub.err; echo "bench sub rc=$?"
CMD="python bench.py --steps 1 --warmup 1 --cpu-seconds 0 --no-e2e --no-side --no-legs"
$CMD > gpurun_out/r2g_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'sga_kernel|sw_classify_kernel|sw_wfa_kernel|sw_fast_kernel' -s 4 -c 4 -o gpurun_out/prof_r2g -f $CMD > gpurun_out/r2g_ncu.log 2>&1
tail -2 gpurun_out/r2g_ncu.log
