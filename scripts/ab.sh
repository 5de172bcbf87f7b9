#!/bin/bash
# A/B of runtime toggles inside ONE gpurun call (same box, same clocks): usage scripts/ab.sh "ENV=.. ENV=.." "ENV=.." ...
for cfg in "$@"; do
  for rep in 1 2; do
    env $cfg python bench.py --steps 3 --warmup 3 --cpu-seconds 0 --no-side ${AB_ARGS} 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); w=d['work']
print('$cfg', 'rep$rep', 'value %.1fM' % (d['value']/1e6), 'ms %.2f' % d['ms_per_step'], 'e2e %.1fM' % (d['e2e']['value']/1e6 if 'e2e' in d else 0), 'stages', {k: round(v,2) for k,v in w['ms_stage'].items()}, 'sga %.2f' % w['ms_sga_kernel'], 'swfast %.2f' % w['ms_sw_fast_kernels'], 'clk', d['clocks'].get('sm_mhz'), 'W', d['clocks'].get('power_w_max'))"
    nvidia-smi --query-gpu=clocks.mem,clocks.sm,memory.used --format=csv,noheader | head -1
  done
done
