#!/bin/bash
# round 2, call 6: 8-lane quick classification, NCCL gather ABI, new bench with legs
python -m pytest tests -m gpu -x -q > gpurun_out/r2f_tests.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/r2f_tests.log
/usr/bin/time -v python bench.py --steps 5 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"
grep -E "Elapsed|Maximum resident" gpurun_out/r2f_bench.err; tail -5 gpurun_out/r2f_bench.err | cut -c1-400
python bench.py --edits sub --steps 5 --warmup 3 --no-legs --cpu-seconds 3 > gpurun_out/r2f_bench_sub.json 2> gpurun_out/r2f_bench_sub.err; echo "bench sub rc=$?"
CMD="python bench.py --steps 1 --warmup 1 --cpu-seconds 0 --no-e2e --no-side --no-legs"
$CMD > gpurun_out/r2f_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'sga_kernel|sw_classify_kernel|sw_wfa_kernel' -s 3 -c 3 -o gpurun_out/prof_r2f -f $CMD > gpurun_out/r2f_ncu.log 2>&1
tail -2 gpurun_out/r2f_ncu.log
