#!/bin/bash
# round 2, call 39: penalty budget of the banded table for reads up to 128 bp (9 / 12 / 15) -- bench A/B, SW parity at 12 and 15
for P in 9 12 15; do
  BWT_SW_BAND_P=$P timeout 300 python bench.py --steps 5 --warmup 3 --no-legs --no-side --cpu-seconds 6 > gpurun_out/r2b8_p$P.json 2> gpurun_out/r2b8_p$P.err
  python - $P <<'PY'
import json, sys
j = [json.loads(l) for l in open("gpurun_out/r2b8_p%s.json" % sys.argv[1]) if l.startswith("{")][-1]
w = j["work"]; e = j["e2e"]
print("budget %s value %.1fM ms %.2f e2e %.1fM sw stage %.2f fast %.2f band frac %.4f shortcut %.4f filled/pair %.0f parity %s/%s" % (sys.argv[1], j["value"] / 1e6, j["ms_per_step"], e["value"] / 1e6,
      w["ms_stage"]["sw"], w["ms_sw_fast_kernels"], w["sw_band_pairs_frac"], w["sw_shortcut_pairs_frac"], w["sw_band_cells_filled_per_pair"], j["parity_sample"]["mismatches"], j["parity_sample"]["candidates"]))
PY
done
for P in 12 15; do
  BWT_SW_BAND_P=$P timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_packed.py tests/test_reference_golden.py -m gpu -x -q -k "sw or align or band or dp_pairs or reference or classification" 2>&1 | tail -1
done
