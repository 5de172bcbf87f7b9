#!/bin/bash
# round 2, call 30 (8 GPUs): e2e at N=8 -- spinning vs blocking waits, 4 vs 2 chunks in flight
run() {
  env $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus 8 --steps 3 --warmup 3 --no-legs --no-side --cpu-seconds 0 $3 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); e=d['e2e']
        print('$1 $3 value %.0fM e2e %.0fM h2d %.1f d2h %.1f GB/s' % (d['value']/1e6, e['value']/1e6, e.get('h2d_gbs',0), e.get('d2h_gbs',0)))"
}
run BWT_BLOCKING_SYNC=0 29541 "--depth 4"
run BWT_BLOCKING_SYNC=1 29542 "--depth 4"
run BWT_BLOCKING_SYNC=0 29543 "--depth 2"
run BWT_BLOCKING_SYNC=1 29544 "--depth 6"
