#!/bin/bash
# round 2, call 18: biased DP keys (packed adds as IMAD on the FMA pipe); sga_kernel occupancy A/B (10 / 11 CTAs per SM)
python -m pytest tests/test_gpu_parity.py tests/test_gpu_packed.py tests/test_gpu_scale.py tests/test_reference_golden.py -m gpu -x -q > gpurun_out/r2r_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2r_tests.log
B="python bench.py --steps 5 --warmup 3 --no-legs --no-side --cpu-seconds 3"
$B > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; echo "bench rc=$?"
python bench.py --config 4 --c4-reads 20000000 --steps 2 --warmup 1 --no-legs --no-side --cpu-seconds 0 > gpurun_out/r2r_c4.json 2> gpurun_out/r2r_c4.err; echo "c4 rc=$?"
BWT_B200_LIB=$PWD/bwt_b200/libbwt_b200_sga10.so $B --no-e2e > gpurun_out/r2r_sga10.json 2> gpurun_out/r2r_sga10.err; echo "sga10 rc=$?"
BWT_B200_LIB=$PWD/bwt_b200/libbwt_b200_sga12.so $B --no-e2e > gpurun_out/r2r_sga12.json 2> gpurun_out/r2r_sga12.err; echo "sga12 rc=$?"
python - <<'P'
import json
for f in ["r2r_bench", "r2r_c4", "r2r_sga10", "r2r_sga12"]:
    try:
        d = [json.loads(l) for l in open("gpurun_out/%s.json" % f) if l.startswith("{")][-1]
        w = d.get("work") or {}
        c4 = d.get("config4") or {}
        print(f, "value %.1fM ms %.2f e2e %.1fM" % (d["value"] / 1e6, d["ms_per_step"], (d.get("e2e") or {}).get("value", 0) / 1e6), "sga", w.get("ms_sga_kernel"), "fast", w.get("ms_sw_fast_kernels"), "stages", w.get("ms_stage"), "parity", (d.get("parity_sample") or {}).get("mismatches"), c4.get("ms_stage_rank0"))
    except Exception as e:
        print(f, "failed", e)
P
