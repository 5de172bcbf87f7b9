#!/bin/bash
# round 2, call 31: sw_band_kernel (banded table + certificate) -- parity in four routings, then A/B on the bench workload
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_packed.py tests/test_reference_golden.py -m gpu -x -q > gpurun_out/r2b1_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2b1_tests.log
for band in 1 0; do
  BWT_SW_BAND=$band timeout 600 python bench.py --steps 5 --warmup 3 --no-legs --no-side --cpu-seconds 8 > gpurun_out/r2b1_bench_band$band.json 2> gpurun_out/r2b1_bench_band$band.err; echo "bench band=$band rc=$?"
done
python - <<'P'
import json
for band in (1, 0):
    try:
        d = [json.loads(l) for l in open("gpurun_out/r2b1_bench_band%d.json" % band) if l.startswith("{")][-1]
        w = d.get("work") or {}
        print("band", band, "value %.1fM ms %.2f e2e %.1fM" % (d["value"] / 1e6, d["ms_per_step"], (d.get("e2e") or {}).get("value", 0) / 1e6),
              "sga", w.get("ms_sga_kernel"), "fast", w.get("ms_sw_fast_kernels"), "stage", w.get("ms_stage"),
              "band frac", w.get("sw_band_pairs_frac"), "shortcut frac", w.get("sw_shortcut_pairs_frac"), "filled/pair", w.get("sw_band_cells_filled_per_pair"),
              "parity", json.dumps(d.get("parity_sample"))[:300])
    except Exception as e:
        print("band", band, "failed", e)
P
