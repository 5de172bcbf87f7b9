#!/bin/bash
# round 2, call 44: ncu --set full of both passes of sw_band_kernel on the final defaults
CMD="python bench.py --steps 1 --warmup 1 --cpu-seconds 0 --no-e2e --no-side --no-legs"
timeout 110 ncu --set full --clock-control none --import-source on -k regex:'sw_band_kernel' -s 2 -c 2 -o gpurun_out/prof_r2fin4 -f $CMD > gpurun_out/r2fin4_ncu.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/r2fin4_ncu.log
