#!/bin/bash
# round 2, call 2: packed reads + fused table + two-stage classification + sync-free run
python -m pytest tests -m gpu -x -q > gpurun_out/r2b_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r2b_tests.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2b_bench_indel.json 2> gpurun_out/r2b_bench_indel.err; echo "bench indel rc=$?"
python bench.py --edits sub --steps 5 --warmup 3 --cpu-seconds 4 > gpurun_out/r2b_bench_sub.json 2> gpurun_out/r2b_bench_sub.err; echo "bench sub rc=$?"
cut -c1-300 gpurun_out/r2b_bench_indel.json; tail -5 gpurun_out/r2b_bench_indel.err
CMD="python bench.py --steps 1 --warmup 1 --cpu-seconds 0 --no-e2e --no-side"
$CMD > gpurun_out/r2b_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'sga_kernel|sw_classify_kernel|sw_wfa_kernel|sw_fast_kernel' -s 4 -c 4 -o gpurun_out/prof_r2b -f $CMD > gpurun_out/r2b_ncu.log 2>&1
tail -2 gpurun_out/r2b_ncu.log; ls -la gpurun_out/*.ncu-rep
