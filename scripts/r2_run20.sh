#!/bin/bash
# round 2, call 20 (2 GPUs): the default line under torchrun; then on one GPU the e2e chunk-size sweep (DP-kernel wave quantisation)
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2t_n2.json 2> gpurun_out/r2t_n2.err; echo "n2 rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --config 5 --steps 2 --warmup 1 > gpurun_out/r2t_n2_c5.json 2> gpurun_out/r2t_n2_c5.err; echo "n2 c5 rc=$?"
python - <<'P'
import json
for f in ["r2t_n2", "r2t_n2_c5"]:
    try:
        d = [json.loads(l) for l in open("gpurun_out/%s.json" % f) if l.startswith("{")][-1]
        print(f, "value %.1fM ms %.2f e2e %.1fM" % (d["value"] / 1e6, d["ms_per_step"], (d.get("e2e") or {}).get("value", 0) / 1e6), "c4", (d.get("config4") or {}).get("value"), "c5", (d.get("config5") or {}).get("value"), (d.get("config5") or {}).get("ms_gather_per_step"), (d.get("config5") or {}).get("nvlink_bytes_per_step"))
    except Exception as e:
        print(f, "failed", e)
P
for cr in 786432 851968 917504 1048576 1310720 1703936 2097152; do
  for dp in 3 4; do
    CUDA_VISIBLE_DEVICES=0 python bench.py --steps 4 --warmup 3 --no-legs --no-side --cpu-seconds 0 --chunk-reads $cr --depth $dp 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); e=d['e2e']
print('chunk $cr depth $dp value %.1fM e2e %.1fM h2d %.2f d2h %.2f ms' % (d['value']/1e6, e['value']/1e6, e['ms_h2d_per_step'], e['ms_d2h_per_step']))"
  done
done
