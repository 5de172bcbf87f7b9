#!/bin/bash
# round 2, call 29: blocking waits for large batches -- N=1 regression check (e2e), spin vs block
B="python bench.py --steps 5 --warmup 3 --no-legs --no-side --cpu-seconds 0"
for m in 0 1 0 1; do
BWT_BLOCKING_SYNC=$m $B 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); e=d['e2e']
print('blocking=$m value %.1fM e2e %.1fM' % (d['value']/1e6, e['value']/1e6))"
done
python -m pytest tests/test_gpu_packed.py tests/test_gpu_scale.py -m gpu -x -q 2>&1 | tail -2
