#!/bin/bash
# N = 2: the default bench line under torchrun (read sharding + the config 4 / 5 legs with the NCCL gather), as the driver launches it
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2j_n2.json 2> gpurun_out/r2j_n2.err; echo "n2 rc=$?"
tail -3 gpurun_out/r2j_n2.err | cut -c1-800
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 1 --impl reference --ref-step-reads 100000 > gpurun_out/r2j_n2_ref.json 2> gpurun_out/r2j_n2_ref.err; echo "n2 ref rc=$?"
cut -c1-300 gpurun_out/r2j_n2_ref.json
