#!/bin/bash
# like ab.sh but with the side measurements and the e2e leg (the allocation pattern of the default bench run)
for cfg in "$@"; do
  env $cfg timeout -s KILL 200 python bench.py --steps 3 --warmup 3 --cpu-seconds 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); w=d['work']
print('$cfg', 'value %.1fM' % (d['value']/1e6), 'ms %.2f' % d['ms_per_step'], 'e2e %.1fM' % (d['e2e']['value']/1e6 if 'e2e' in d else 0), 'stages', {k: round(v,2) for k,v in w['ms_stage'].items()}, 'sga %.2f' % w['ms_sga_kernel'], 'swfast %.2f' % w['ms_sw_fast_kernels'])"
done
