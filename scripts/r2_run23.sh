#!/bin/bash
# round 2, call 23: register sort / vote in the general seeding path (reads with 'N')
python -m pytest tests -m gpu -x -q > gpurun_out/r2w_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2w_tests.log
python scripts/n_reads_probe.py 2>&1 | tail -1
python bench.py --steps 5 --warmup 3 --no-side --cpu-seconds 3 > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; echo "bench rc=$?"
python - <<'P'
import json
d = [json.loads(l) for l in open("gpurun_out/r2w_bench.json") if l.startswith("{")][-1]
w = d["work"]
print("value %.1fM ms %.2f e2e %.1fM sga %.2f fast %.2f parity %s" % (d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6, w["ms_sga_kernel"], w["ms_sw_fast_kernels"], d["parity_sample"]["mismatches"]))
print("n_reads", json.dumps(d["n_reads"])[:700])
print("config1", json.dumps(d["config1"])[:400])
print("config4", d["config4"]["value"], d["config4"]["e2e"]["value"], "config5", d["config5"]["value"])
P
