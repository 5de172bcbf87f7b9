#!/bin/bash
# round 2, call 32: where the e2e leg loses against the resident step (per-chunk host timeline), and sw_band_kernel under ncu
BWT_TRACE=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-legs --no-side --cpu-seconds 0 > gpurun_out/r2b2_bench_trace.json 2> gpurun_out/r2b2_trace.err; echo "trace rc=$?"
grep -c "^chunk" gpurun_out/r2b2_trace.err
CMD="python bench.py --steps 1 --warmup 1 --cpu-seconds 0 --no-e2e --no-side --no-legs"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b2_launches.csv $CMD > gpurun_out/r2b2_ncu_launch.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'sw_band_kernel' -s 1 -c 1 -o gpurun_out/prof_r2b2 -f $CMD > gpurun_out/r2b2_ncu.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/r2b2_ncu.log
