#!/bin/bash
# round 2, call 21 (8 GPUs): config 3 weak scaling and config 5 with one partition per GPU, as the driver launches them
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 3 --warmup 3 --no-legs --no-side --cpu-seconds 0 > gpurun_out/r2u_n8.json 2> gpurun_out/r2u_n8.err; echo "n8 rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --config 5 --steps 2 --warmup 1 > gpurun_out/r2u_n8_c5.json 2> gpurun_out/r2u_n8_c5.err; echo "n8 c5 rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --config 4 --steps 2 --warmup 1 > gpurun_out/r2u_n8_c4.json 2> gpurun_out/r2u_n8_c4.err; echo "n8 c4 rc=$?"
nvidia-smi topo -m > gpurun_out/r2u_topo.txt 2>&1; nproc >> gpurun_out/r2u_topo.txt; lscpu | grep -i "numa\|socket\|model name" >> gpurun_out/r2u_topo.txt
python - <<'P'
import json
for f in ["r2u_n8", "r2u_n8_c5", "r2u_n8_c4"]:
    try:
        d = [json.loads(l) for l in open("gpurun_out/%s.json" % f) if l.startswith("{")][-1]
        e = d.get("e2e") or {}
        print(f, "value %.1fM ms %.2f e2e %.1fM h2d %s d2h %s" % (d["value"] / 1e6, d["ms_per_step"], e.get("value", 0) / 1e6, e.get("h2d_gbs"), e.get("d2h_gbs")), "c5", json.dumps(d.get("config5") or {})[:600] if "c5" in f else "", "c4", json.dumps(d.get("config4") or {})[:500] if "c4" in f else "")
    except Exception as e:
        print(f, "failed", e)
P
tail -3 gpurun_out/r2u_n8.err | cut -c1-300
