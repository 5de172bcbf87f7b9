#!/bin/bash
# ncu evidence, second capture of round 1 (deep k-mer table, packed SW, register-sort sga)
CMD="python bench.py --reads 2000000 --steps 1 --warmup 1 --cpu-seconds 0 --no-e2e"
$CMD > gpurun_out/plain_b.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_b.csv $CMD > gpurun_out/ncu1b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'sga_kernel|sw_fast_kernel' -c 3 -o gpurun_out/prof_r1b -f $CMD > gpurun_out/ncu2b.log 2>&1
tail -1 gpurun_out/plain_b.log | cut -c1-300; tail -3 gpurun_out/ncu2b.log
