#!/bin/bash
# full GPU check: engine + parity tests, smoke, small + mid bench, launch list
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
nproc; free -g | head -2
echo "=== gpu tests"; timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -30
echo "=== smoke"; timeout 600 python __graft_entry__.py smoke 2>&1 | tail -12
echo "=== bench cfg5 n=262144"; timeout 900 python bench.py --n 262144 --steps 3 --warmup 3 2>&1 | tail -3
echo "=== bench cfg3 n=262144"; timeout 900 python bench.py --workload cfg3 --n 262144 --steps 3 --warmup 3 2>&1 | tail -3
echo "=== bench cfg4 n=131072"; timeout 900 python bench.py --workload cfg4 --n 131072 --steps 3 --warmup 3 2>&1 | tail -3
} > gpurun_out/run4.log 2>&1
tail -40 gpurun_out/run4.log
