#!/bin/bash
python scripts/overlap_probe.py > gpurun_out/r2h_overlap.log 2>&1; echo "overlap rc=$?"; cat gpurun_out/r2h_overlap.log
for cr in 524288 1048576 2097152; do for d in 2 3 4; do
python bench.py --steps 4 --warmup 2 --no-legs --no-side --cpu-seconds 0 --chunk-reads $cr --depth $d 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunk $cr depth $d value %.1fM e2e %.1fM e2e_ascii %.1fM' % (d['value']/1e6, d['e2e']['value']/1e6, d['e2e_ascii']['value']/1e6))"
done; done 2>&1 | tee gpurun_out/r2h_sweep.log
