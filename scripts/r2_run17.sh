#!/bin/bash
# round 2, call 17: row maxima of the strip's last row only when the slack allows (sw_fast_kernel)
python -m pytest tests/test_gpu_parity.py tests/test_gpu_packed.py tests/test_gpu_scale.py -m gpu -x -q > gpurun_out/r2q_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2q_tests.log
B="python bench.py --steps 5 --warmup 3 --no-legs --no-side --cpu-seconds 3"
$B > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; echo "bench rc=$?"
python bench.py --config 4 --c4-reads 20000000 --steps 2 --warmup 1 --no-legs --no-side --cpu-seconds 0 > gpurun_out/r2q_c4.json 2> gpurun_out/r2q_c4.err; echo "c4 rc=$?"
python - <<'P'
import json
for f in ["r2q_bench", "r2q_c4"]:
    try:
        d = [json.loads(l) for l in open("gpurun_out/%s.json" % f) if l.startswith("{")][-1]
        w = d.get("work") or {}
        c4 = d.get("config4") or {}
        print(f, "value %.1fM ms %.2f e2e %.1fM" % (d["value"] / 1e6, d["ms_per_step"], (d.get("e2e") or {}).get("value", 0) / 1e6), "sga", w.get("ms_sga_kernel"), "fast", w.get("ms_sw_fast_kernels"), "stages", w.get("ms_stage"), "parity", (d.get("parity_sample") or {}).get("mismatches"), c4.get("ms_stage_rank0"))
    except Exception as e:
        print(f, "failed", e)
P
python scripts/n_reads_probe.py 2>&1 | tail -1
