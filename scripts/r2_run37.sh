#!/bin/bash
# round 2, call 37: chunk ramp A/B on the e2e leg, packed / scale tests on the new aligner defaults
for ramp in 1 0 1 0; do
  BWT_CHUNK_RAMP=$ramp timeout 300 python bench.py --steps 4 --warmup 3 --no-legs --no-side --cpu-seconds 0 > gpurun_out/r2b7_ramp$ramp.json 2> gpurun_out/r2b7_ramp$ramp.err
  python - $ramp <<'P'
import json, sys
j = [json.loads(l) for l in open("gpurun_out/r2b7_ramp%s.json" % sys.argv[1]) if l.startswith("{")][-1]
e = j["e2e"]
print("ramp %s value %.1fM e2e %.1fM h2d %.2f d2h %.2f ms depth %s" % (sys.argv[1], j["value"] / 1e6, e["value"] / 1e6, e["ms_h2d_per_step"], e["ms_d2h_per_step"], e.get("chunks_in_flight")))
P
done
timeout 600 python -m pytest tests/test_gpu_packed.py tests/test_gpu_scale.py tests/test_gpu_nif_paths.py -m gpu -x -q 2>&1 | tail -2
