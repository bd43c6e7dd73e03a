#!/bin/bash
# 2 GPUs: sharded parity (dense fused fit, dense unfused, sparse) and the scaling bench line
mkdir -p gpurun_out
{
nvidia-smi -L
echo "=== 2-rank parity"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_parity.py 2>&1 | grep -v "^\*\|OMP_NUM" | tail -16
echo "=== bench --gpus 2"
T0=$(date +%s)
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 2>&1 | tail -1
echo "wall $(( $(date +%s) - T0 )) s"
echo "=== bench --gpus 2 --impl reference"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-300
} > gpurun_out/run27.log 2>&1
cat gpurun_out/run27.log | cut -c1-2500
