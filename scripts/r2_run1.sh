#!/bin/bash
# round 2, call 1: r1 kernels on the new (indel) workload + in-run parity sample; source-level ncu of the two top kernels
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2a_tests.log
python bench.py --edits sub --steps 5 --warmup 3 > gpurun_out/r2a_bench_sub.json 2> gpurun_out/r2a_bench_sub.err; echo "bench sub rc=$?"
python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench_indel.json 2> gpurun_out/r2a_bench_indel.err; echo "bench indel rc=$?"
CMD="python bench.py --steps 1 --warmup 1 --cpu-seconds 0 --no-e2e --no-side"
$CMD > gpurun_out/r2a_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'sga_kernel|sw_classify_kernel' -s 2 -c 2 -o gpurun_out/prof_r2a -f $CMD > gpurun_out/r2a_ncu.log 2>&1
tail -2 gpurun_out/r2a_ncu.log; ls -la gpurun_out/*.ncu-rep
cut -c1-600 gpurun_out/r2a_bench_indel.json; tail -5 gpurun_out/r2a_bench_indel.err
