#!/bin/bash
for cfg in "524288 3" "524288 4" "1048576 4" "2097152 3" "262144 6" "1048576 2"; do
  set -- $cfg
  echo "chunk=$1 depth=$2: $(python bench.py --cpu-seconds 0 --steps 3 --warmup 1 --chunk-reads $1 --depth $2 2>&1 | tail -1 | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("value %.1fM e2e %.1fM" % (d["value"]/1e6, d["e2e"]["value"]/1e6))')"
done
