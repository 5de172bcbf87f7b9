#!/bin/bash
# round 2, call 28: parity at FULL bench size -- the CPU port's candidates for all 10 M reads of a step against the GPU's
python bench.py --steps 3 --warmup 3 --no-legs --no-side --cpu-seconds 75 > gpurun_out/r2p10m_bench.json 2> gpurun_out/r2p10m_bench.err; echo "bench rc=$?"
python - <<'P'
import json
d = [json.loads(l) for l in open("gpurun_out/r2p10m_bench.json") if l.startswith("{")][-1]
print("value %.1fM" % (d["value"] / 1e6), json.dumps(d["parity_sample"]), json.dumps(d["cpu_baseline"])[:400])
P
python -m pytest tests/test_reference_golden.py -m gpu -x -q 2>&1 | tail -2
