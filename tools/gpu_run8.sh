#!/bin/bash
# 2-GPU check: the driver's torchrun launch of the default bench + a 2-rank parity test of the sharded fit
mkdir -p gpurun_out
{
nvidia-smi -L
echo "=== 2-rank parity"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_parity.py 2>&1 | tail -8
echo "=== bench --gpus 2"; S=$(date +%s); timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 2>&1 | tail -4; echo "wall $(( $(date +%s) - S )) s"
} > gpurun_out/run8.log 2>&1
exit 0
