#!/bin/bash
# ncu evidence for round 1: launch list + full capture of the two dominant kernels (run under gpurun, 1 GPU)
set -x
CMD="python bench.py --reads 2000000 --steps 1 --warmup 1 --cpu-seconds 0 --no-e2e"
python -c "
import sys; sys.path.insert(0,'.')
from bwt_b200 import api
print('gather_gbs_8GiB', api.peak_gather_gbs(0, 8<<30))
print('gather_gbs_1GiB', api.peak_gather_gbs(0, 1<<30))
print('dpx_gops', api.peak_dpx_gops(0))
" > gpurun_out/peaks.txt 2>&1
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'sga_kernel|sw_fast_kernel' -c 4 -o gpurun_out/prof_r1 -f $CMD > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/plain.log; cat gpurun_out/peaks.txt; tail -3 gpurun_out/ncu2.log
