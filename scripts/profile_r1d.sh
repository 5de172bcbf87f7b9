#!/bin/bash
# ncu evidence at the bench's own workload (10 M reads per launch) after the sw shortcut; run under gpurun on 1 GPU
CMD="python bench.py --steps 1 --warmup 1 --cpu-seconds 0 --no-e2e --no-side"
$CMD > gpurun_out/plain_d.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_d.csv $CMD > gpurun_out/ncu1d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'sga_kernel|sw_fast_kernel|sw_classify_kernel' -c 6 -o gpurun_out/prof_r1d -f $CMD > gpurun_out/ncu2d.log 2>&1
tail -1 gpurun_out/plain_d.log | cut -c1-200; tail -2 gpurun_out/ncu2d.log
# the bounded-edit search (inexact.cu): launch list + full capture of fs_search_kernel
CMD2="python scripts/bench_inexact.py --ref-bp 400000000 --reads 100000 --steps 1 --warmup 1"
$CMD2 > gpurun_out/plain_fs.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'fs_search_kernel' -c 2 -o gpurun_out/prof_r1d_fs -f $CMD2 > gpurun_out/ncu_fs.log 2>&1
tail -1 gpurun_out/plain_fs.log | cut -c1-300; tail -2 gpurun_out/ncu_fs.log
