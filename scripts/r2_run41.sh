#!/bin/bash
# round 2, call 41: the default bench line with every leg on the final defaults (two band passes, ordered aligner)
timeout 400 python bench.py > gpurun_out/r2fin4_bench.json 2> gpurun_out/r2fin4_bench.err; echo "bench rc=$?"
python - <<'P'
import json
d = [json.loads(l) for l in open("gpurun_out/r2fin4_bench.json") if l.startswith("{")][-1]
w = d["work"]
print("value %.1fM ms %.2f e2e %.1fM" % (d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6), "stage", w["ms_stage"], "parity", d["parity_sample"]["mismatches"], d["parity_sample"]["candidates"])
print("config4 %.1fM e2e %.1fM" % (d["config4"]["value"] / 1e6, d["config4"]["e2e"]["value"] / 1e6), "config5 %.2fM" % (d["config5"]["value"] / 1e6), "config1 ms/call", d["config1"]["ms_per_200_read_call"], "n_reads %.1fM" % (d["n_reads"]["value"] / 1e6), "clocks", d["clocks"])
P
