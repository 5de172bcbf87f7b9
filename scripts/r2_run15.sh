#!/bin/bash
# round 2, call 15: bit-parallel window verification in the bounded-edit front end; FP64 probe
python -m pytest tests/test_gpu_inexact.py tests/test_gpu_parity.py tests/test_reference_golden.py -m gpu -x -q > gpurun_out/r2o_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2o_tests.log
./scripts/probes/fp64_probe > gpurun_out/r2o_fp64.log 2>&1; cat gpurun_out/r2o_fp64.log
python bench.py --steps 5 --warmup 3 --cpu-seconds 3 > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; echo "bench rc=$?"
python - <<'P'
import json
d = [json.loads(l) for l in open("gpurun_out/r2o_bench.json") if l.startswith("{")][-1]
w = d["work"]
print("value %.1fM ms %.2f sga %.2f fast %.2f e2e %.1fM parity %s" % (d["value"] / 1e6, d["ms_per_step"], w["ms_sga_kernel"], w["ms_sw_fast_kernels"], d["e2e"]["value"] / 1e6, d["parity_sample"]["mismatches"]))
print("inexact", json.dumps(d.get("inexact"))[:600])
print("config5", json.dumps((d.get("config5") or {}))[:1500])
print("roofline", json.dumps(d["roofline"])[:900])
P
