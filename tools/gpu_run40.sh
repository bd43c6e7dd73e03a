#!/bin/bash
# N-GPU scaling points (N = $1): cfg5 strong scaling (default workload), cfg4 full size, cfg3 full size
N=${1:-8}
mkdir -p gpurun_out
{
nvidia-smi -L | head -8
for w in cfg5 cfg4 cfg3; do
echo "=== bench --gpus $N --workload $w"; S=$(date +%s)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --workload $w --no-cpu --alt-mode= 2>&1 | tail -2
echo "wall $(( $(date +%s) - S )) s"
done
} > gpurun_out/run40_${N}gpu.log 2>&1
cut -c1-2500 gpurun_out/run40_${N}gpu.log
