#!/bin/bash
# 2-GPU line of bench.py (torchrun, one rank per GPU): run with gpurun --gpus 2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/scale2.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --workload 1080p --steps 3 --warmup 3 > gpurun_out/scale2_1080p.json 2>> gpurun_out/scale2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/scale2_2160p.json 2>> gpurun_out/scale2.log
tail -5 gpurun_out/scale2.log; cat gpurun_out/scale2_1080p.json gpurun_out/scale2_2160p.json | cut -c1-900
