#!/bin/bash
# round 2, call 25: e2e depth sweep (chunks in flight)
for cr in 524288 1048576; do
  for dp in 4 6 8; do
    python bench.py --steps 4 --warmup 3 --no-legs --no-side --cpu-seconds 0 --chunk-reads $cr --depth $dp 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); e=d['e2e']
print('chunk $cr depth $dp value %.1fM e2e %.1fM h2d %.2f d2h %.2f ms' % (d['value']/1e6, e['value']/1e6, e['ms_h2d_per_step'], e['ms_d2h_per_step']))"
  done
done
