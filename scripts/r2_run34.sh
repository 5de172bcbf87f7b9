#!/bin/bash
# round 2, call 34: chunks inside their run at a time (BWT_COMPUTE_TOKEN) x depth x chunk size, e2e leg
run() {  # token depth chunk
  BWT_COMPUTE_TOKEN=$1 timeout 300 python bench.py --steps 4 --warmup 3 --no-legs --no-side --cpu-seconds 0 --depth $2 --chunk-reads $3 > gpurun_out/r2b4_t$1_d$2_c$3.json 2> gpurun_out/r2b4_t$1_d$2_c$3.err
  python - "$1" "$2" "$3" <<'P'
import json, sys
t, d, c = sys.argv[1:4]
try:
    j = [json.loads(l) for l in open("gpurun_out/r2b4_t%s_d%s_c%s.json" % (t, d, c)) if l.startswith("{")][-1]
    e = j["e2e"]
    print("token %s depth %s chunk %s value %.1fM e2e %.1fM h2d %.2f d2h %.2f ms" % (t, d, c, j["value"] / 1e6, e["value"] / 1e6, e["ms_h2d_per_step"], e["ms_d2h_per_step"]))
except Exception as ex:
    print("token %s depth %s chunk %s failed %s" % (t, d, c, ex))
P
}
run 0 4 1048576
run 1 4 1048576
run 2 4 1048576
run 3 4 1048576
run 2 3 1048576
run 2 6 1048576
run 2 6 524288
run 3 6 524288
timeout 300 python -m pytest tests/test_gpu_packed.py tests/test_gpu_scale.py -m gpu -x -q 2>&1 | tail -2
