#!/bin/bash
# round 2, call 40: banded table -- list ordered by shape in front of it or behind it, one pass (budget 15) or two (9, then 15 on the rest);
# then the full GPU suite, the bench line and the launch list on the defaults (two passes, ordered in front)
run() {  # tag env...
  tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-legs --no-side --cpu-seconds 4 > gpurun_out/r2b9_$tag.json 2> gpurun_out/r2b9_$tag.err
  python - $tag "$*" <<'PY'
import json, sys
j = [json.loads(l) for l in open("gpurun_out/r2b9_%s.json" % sys.argv[1]) if l.startswith("{")][-1]
w = j["work"]; e = j["e2e"]
print("%-28s value %.1fM ms %.2f e2e %.1fM sw stage %.2f fast %.2f band frac %.4f filled/pair %.0f launches %s parity %s/%s" % (sys.argv[2], j["value"] / 1e6, j["ms_per_step"], e["value"] / 1e6,
      w["ms_stage"]["sw"], w["ms_sw_fast_kernels"], w["sw_band_pairs_frac"], w["sw_band_cells_filled_per_pair"], j["gpu_launches"], j["parity_sample"]["mismatches"], j["parity_sample"]["candidates"]))
PY
}
run default BWT_X=0
run p15 BWT_SW_BAND_P=15 BWT_SW_BAND_P2=0 BWT_SW_BAND_SORT=0
run p15s BWT_SW_BAND_P=15 BWT_SW_BAND_P2=0 BWT_SW_BAND_SORT=1
run p9p15u BWT_SW_BAND_P=9 BWT_SW_BAND_P2=15 BWT_SW_BAND_SORT=0
run p9s BWT_SW_BAND_P=9 BWT_SW_BAND_P2=0 BWT_SW_BAND_SORT=1
python -m pytest tests -m gpu -x -q > gpurun_out/r2fin4_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2fin4_tests.log
CMD="python bench.py --steps 1 --warmup 1 --cpu-seconds 0 --no-e2e --no-side --no-legs"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2fin4_launches.csv $CMD > gpurun_out/r2fin4_ncu_launch.log 2>&1; echo "launch list rc=$?"
