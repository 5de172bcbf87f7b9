#!/bin/bash
# ncu evidence at the bench's own workload (10 M reads per launch); run under gpurun on 1 GPU
CMD="python bench.py --steps 1 --warmup 1 --cpu-seconds 0 --no-e2e --no-side"
$CMD > gpurun_out/plain_c.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c.csv $CMD > gpurun_out/ncu1c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'sga_kernel|sw_fast_kernel' -c 4 -o gpurun_out/prof_r1c -f $CMD > gpurun_out/ncu2c.log 2>&1
tail -1 gpurun_out/plain_c.log | cut -c1-200; tail -2 gpurun_out/ncu2c.log
