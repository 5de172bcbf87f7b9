#!/bin/bash
# round 2, call 16: fewer ALU-pipe instructions per DP cell; occupancy A/B of the warp-per-pair kernel on the 'N' workload
python -m pytest tests/test_gpu_parity.py tests/test_gpu_packed.py tests/test_gpu_scale.py -m gpu -x -q > gpurun_out/r2p_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2p_tests.log
B="python bench.py --steps 5 --warmup 3 --no-legs --no-side --cpu-seconds 3"
$B > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; echo "bench rc=$?"
$B --edits sub > gpurun_out/r2p_bench_sub.json 2> gpurun_out/r2p_bench_sub.err; echo "bench sub rc=$?"
python bench.py --config 4 --c4-reads 20000000 --steps 2 --warmup 1 --no-legs --no-side --cpu-seconds 0 > gpurun_out/r2p_c4.json 2> gpurun_out/r2p_c4.err; echo "c4 rc=$?"
python - <<'P'
import json
for f in ["r2p_bench", "r2p_bench_sub", "r2p_c4"]:
    try:
        d = [json.loads(l) for l in open("gpurun_out/%s.json" % f) if l.startswith("{")][-1]
        w = d.get("work") or {}
        c4 = d.get("config4") or {}
        print(f, "value %.1fM ms %.2f e2e %.1fM" % (d["value"] / 1e6, d["ms_per_step"], (d.get("e2e") or {}).get("value", 0) / 1e6), "sga", w.get("ms_sga_kernel"), "fast", w.get("ms_sw_fast_kernels"), "stages", w.get("ms_stage"), "parity", (d.get("parity_sample") or {}).get("mismatches"), c4.get("ms_stage_rank0"))
    except Exception as e:
        print(f, "failed", e)
P
python scripts/n_reads_probe.py 2>&1 | tail -1
BWT_B200_LIB=$PWD/bwt_b200/libbwt_b200_gw8.so python scripts/n_reads_probe.py 2>&1 | tail -1
BWT_B200_LIB=$PWD/bwt_b200/libbwt_b200_gw12.so python scripts/n_reads_probe.py 2>&1 | tail -1
