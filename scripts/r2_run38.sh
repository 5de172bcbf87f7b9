#!/bin/bash
# round 2, call 38 (8 GPUs): config 3 at N=8 with the band kernel and the ordered aligner; chunks in flight auto (by host cores) vs 5
run() {
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 --steps 3 --warmup 3 --no-legs --no-side --cpu-seconds 0 $2 2>/dev/null > gpurun_out/r2n8c_$1.json
  python - $1 "$2" <<'P'
import json, sys
for l in open("gpurun_out/r2n8c_%s.json" % sys.argv[1]):
    if l.startswith('{'):
        d = json.loads(l); e = d['e2e']
        print('%s value %.0fM e2e %.0fM depth %s h2d %.1f d2h %.1f GB/s' % (sys.argv[2], d['value'] / 1e6, e['value'] / 1e6, e.get('chunks_in_flight'), e.get('h2d_gbs', 0), e.get('d2h_gbs', 0)))
P
}
nproc
run 29551 ""
run 29552 "--depth 5"
