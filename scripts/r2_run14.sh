#!/bin/bash
# round 2, call 14: restructured warp-per-pair kernel (row chain, cooperative traceback), sga with 8 lanes per read, L2::64B loads
python -m pytest tests -m gpu -x -q > gpurun_out/r2n_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2n_tests.log
python scripts/small_call_probe.py > gpurun_out/r2n_small.log 2>&1; echo "small rc=$?"; cat gpurun_out/r2n_small.log
B="python bench.py --steps 5 --warmup 3 --no-legs --no-side --no-e2e --cpu-seconds 3"
BWT_SGA_G=8 $B > gpurun_out/r2n_g8.json 2> gpurun_out/r2n_g8.err; echo "g8 rc=$?"
BWT_SGA_PREFETCH=8 $B > gpurun_out/r2n_h64.json 2> gpurun_out/r2n_h64.err; echo "h64 rc=$?"
BWT_SGA_G=8 BWT_SGA_PREFETCH=8 $B > gpurun_out/r2n_g8h64.json 2> gpurun_out/r2n_g8h64.err; echo "g8h64 rc=$?"
python bench.py --steps 5 --warmup 3 --cpu-seconds 3 > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; echo "bench rc=$?"
python - <<'P'
import json
for f in ["r2n_g8", "r2n_h64", "r2n_g8h64", "r2n_bench"]:
    try:
        d = [json.loads(l) for l in open("gpurun_out/%s.json" % f) if l.startswith("{")][-1]
        w = d["work"]
        print(f, "value %.1fM ms %.2f sga %.2f fast %.2f parity %s" % (d["value"] / 1e6, d["ms_per_step"], w["ms_sga_kernel"], w["ms_sw_fast_kernels"], d.get("parity_sample", {}).get("mismatches")),
              "e2e %.1fM" % (d.get("e2e", {}).get("value", 0) / 1e6), "n_reads", (d.get("n_reads") or {}).get("value"), (d.get("n_reads") or {}).get("ms_stage"), "c1", (d.get("config1") or {}).get("ms_per_200_read_call"))
    except Exception as e:
        print(f, "failed", e)
P
CMD="python bench.py --steps 1 --warmup 1 --cpu-seconds 0 --no-e2e --no-side --no-legs"
BWT_SGA_PREFETCH=8 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:sga_kernel -s 1 -c 1 --csv --log-file gpurun_out/r2n_sga_h64.csv $CMD > /dev/null 2>&1; echo "ncu h64 rc=$?"; tail -3 gpurun_out/r2n_sga_h64.csv | cut -d, -f13-
python - <<'P'
import json
d = [json.loads(l) for l in open("gpurun_out/r2n_bench.json") if l.startswith("{")][-1]
print("inexact", json.dumps(d.get("inexact"))[:700])
print("config5 inexact", json.dumps((d.get("config5") or {}).get("inexact"))[:500])
P
