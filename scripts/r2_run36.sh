#!/bin/bash
# round 2, call 36: e2e leg, chunks inside their run x chunks in flight x chunk size (second sweep)
run() {  # token depth chunk
  BWT_COMPUTE_TOKEN=$1 timeout 300 python bench.py --steps 4 --warmup 3 --no-legs --no-side --cpu-seconds 0 --depth $2 --chunk-reads $3 > gpurun_out/r2b6_t$1_d$2_c$3.json 2> gpurun_out/r2b6_t$1_d$2_c$3.err
  python - "$1" "$2" "$3" <<'P'
import json, sys
t, d, c = sys.argv[1:4]
try:
    j = [json.loads(l) for l in open("gpurun_out/r2b6_t%s_d%s_c%s.json" % (t, d, c)) if l.startswith("{")][-1]
    e = j["e2e"]
    print("token %s depth %s chunk %s value %.1fM e2e %.1fM h2d %.2f d2h %.2f ms" % (t, d, c, j["value"] / 1e6, e["value"] / 1e6, e["ms_h2d_per_step"], e["ms_d2h_per_step"]))
except Exception as ex:
    print("token %s depth %s chunk %s failed %s" % (t, d, c, ex))
P
}
run 3 5 1048576
run 3 6 1048576
run 3 8 1048576
run 4 6 1048576
run 4 8 1048576
run 5 8 1048576
run 3 5 1572864
run 3 6 1572864
run 4 6 1572864
run 3 5 2097152
run 3 6 2097152
run 4 6 2097152
run 3 6 786432
run 4 8 786432
