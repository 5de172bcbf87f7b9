#!/bin/bash
# round 2, call 24: 16 lanes per read for 150-base reads (config 4) A/B; ncu of the double-precision and bounded-edit kernels
C4="python bench.py --config 4 --c4-reads 20000000 --steps 2 --warmup 1 --no-legs --no-side --cpu-seconds 0"
BWT_SGA_G16_MAXLEN=0 $C4 > gpurun_out/r2x_c4_g32.json 2> gpurun_out/r2x_c4_g32.err; echo "c4 g32 rc=$?"
$C4 > gpurun_out/r2x_c4_g16.json 2> gpurun_out/r2x_c4_g16.err; echo "c4 g16 rc=$?"
python - <<'P'
import json
for f in ["r2x_c4_g32", "r2x_c4_g16"]:
    try:
        d = [json.loads(l) for l in open("gpurun_out/%s.json" % f) if l.startswith("{")][-1]
        c4 = d.get("config4") or {}
        print(f, "value %.1fM ms %.2f e2e %.1fM" % (d["value"] / 1e6, d["ms_per_step"], (d.get("e2e") or {}).get("value", 0) / 1e6), c4.get("ms_stage_rank0"), c4.get("hits_per_step"))
    except Exception as e:
        print(f, "failed", e)
P
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "150 or long or sga" > gpurun_out/r2x_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2x_tests.log
ncu --set full --clock-control none --import-source on -k regex:'sw_double_kernel|sw_general_warp_kernel|sga_kernel' -c 6 -o gpurun_out/prof_r2x_nreads -f python scripts/n_reads_probe.py > gpurun_out/r2x_ncu_nreads.log 2>&1; echo "ncu nreads rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'fs_search_kernel' -c 2 -o gpurun_out/prof_r2x_fs -f python scripts/bench_inexact.py --reads 50000 --steps 1 --warmup 1 > gpurun_out/r2x_ncu_fs.log 2>&1; echo "ncu fs rc=$?"; tail -3 gpurun_out/r2x_ncu_fs.log
