#!/bin/bash
# round 2, call 33: band kernel with grouped string loads / batched walk; e2e timeline with device-side times, fully warmed
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_packed.py tests/test_reference_golden.py -m gpu -x -q > gpurun_out/r2b3_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2b3_tests.log
BWT_TRACE=1 timeout 600 python bench.py --steps 4 --warmup 3 --no-legs --no-side --cpu-seconds 0 > gpurun_out/r2b3_bench.json 2> gpurun_out/r2b3_trace.err; echo "bench rc=$?"
python - <<'P'
import json
d = [json.loads(l) for l in open("gpurun_out/r2b3_bench.json") if l.startswith("{")][-1]
w = d.get("work") or {}
print("value %.1fM ms %.2f e2e %.1fM" % (d["value"] / 1e6, d["ms_per_step"], (d.get("e2e") or {}).get("value", 0) / 1e6), "sga", w.get("ms_sga_kernel"), "fast", w.get("ms_sw_fast_kernels"), "stage", w.get("ms_stage"), json.dumps(d.get("e2e")))
P
