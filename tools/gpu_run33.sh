#!/bin/bash
# session 4, run 1: full gpu test suite on the restored tree, cfg3 with the k=256 cluster kernel on / off,
# cfg4 at n=500000 with the blocked-CSC numerator, ncu of the k=256 fused kernel
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,power.limit --format=csv
echo "=== gpu tests"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
echo "=== cfg3 fused256 on";  timeout 600 python bench.py --workload cfg3 --no-cpu --alt-mode= 2>&1 | tail -2
echo "=== cfg3 fused256 off"; KLNMF_FUSED256=0 timeout 600 python bench.py --workload cfg3 --no-cpu --no-e2e --alt-mode= 2>&1 | tail -2
echo "=== cfg3 k=192 on"; timeout 600 python bench.py --workload cfg3 --k 192 --no-cpu --no-e2e --alt-mode= 2>&1 | tail -1
echo "=== cfg3 k=192 off"; KLNMF_FUSED256=0 timeout 600 python bench.py --workload cfg3 --k 192 --no-cpu --no-e2e --alt-mode= 2>&1 | tail -1
echo "=== cfg4 n=500000"; timeout 600 python bench.py --workload cfg4 --n 500000 --no-cpu --alt-mode= 2>&1 | tail -2
echo "=== ncu full fused256"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_coef256 -s 1 -c 2 -f -o gpurun_out/r1_fused256_cfg3_n262144 python bench.py --workload cfg3 --n 262144 --steps 2 --warmup 1 --no-cpu --no-e2e --alt-mode= 2>&1 | tail -3
} > gpurun_out/run33.log 2>&1
cut -c1-1500 gpurun_out/run33.log
