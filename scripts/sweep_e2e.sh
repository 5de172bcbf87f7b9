#!/bin/bash
# e2e (host buffers -> hit tuples) against chunk size and chunks in flight of the pipelined aligner
for cfg in "1048576 2" "1048576 3" "1048576 4" "524288 4" "524288 6" "2097152 3" "262144 8"; do
  set -- $cfg
  echo "chunk=$1 depth=$2: $(timeout -s KILL 120 python bench.py --cpu-seconds 0 --no-side --steps 3 --warmup 2 --chunk-reads $1 --depth $2 2>/dev/null | tail -1 | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("value %.1fM e2e %.1fM" % (d["value"]/1e6, d["e2e"]["value"]/1e6))')"
done
