#!/bin/bash
# round 2, call 35: e2e timeline with two / three chunks inside their run at a time
for t in 2 3; do
  BWT_TRACE=1 BWT_COMPUTE_TOKEN=$t timeout 300 python bench.py --steps 2 --warmup 3 --no-legs --no-side --cpu-seconds 0 > gpurun_out/r2b5_t$t.json 2> gpurun_out/r2b5_trace_t$t.err; echo "token $t rc=$?"
done
