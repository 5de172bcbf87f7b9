#!/bin/bash
# round 2, call 22: thread-per-pair double-precision kernel (swd.cu); DP list largest-first
python -m pytest tests -m gpu -x -q > gpurun_out/r2v_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r2v_tests.log
python scripts/n_reads_probe.py 2>&1 | tail -1
BWT_B200_LIB=$PWD/bwt_b200/libbwt_b200_swd3.so python scripts/n_reads_probe.py 2>&1 | tail -1
BWT_B200_LIB=$PWD/bwt_b200/libbwt_b200_swd5.so python scripts/n_reads_probe.py 2>&1 | tail -1
BWT_SW_DOUBLE=0 python scripts/n_reads_probe.py 2>&1 | tail -1
python bench.py --steps 5 --warmup 3 --no-legs --no-side --cpu-seconds 3 > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; echo "bench rc=$?"
python - <<'P'
import json
d = [json.loads(l) for l in open("gpurun_out/r2v_bench.json") if l.startswith("{")][-1]
w = d["work"]
print("value %.1fM ms %.2f e2e %.1fM sga %.2f fast %.2f parity %s" % (d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6, w["ms_sga_kernel"], w["ms_sw_fast_kernels"], d["parity_sample"]["mismatches"]))
P
