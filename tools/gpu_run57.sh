#!/bin/bash
# 2 GPUs: sharded parity with the centered split-TF32 ratio (colsum(W') joins the all-reduced doubles), short bench in both modes
mkdir -p gpurun_out
{
nvidia-smi -L
echo "=== 2-rank parity"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_parity.py 2>&1 | grep -v "^\*\|OMP_NUM" | tail -16
echo "=== bench --gpus 2 (n=1M, tf32 + tf32x3)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --n 1000000 --steps 5 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | cut -c1-2500
} > gpurun_out/run57.log 2>&1
cut -c1-2500 gpurun_out/run57.log
