#!/bin/bash
mkdir -p gpurun_out
{
echo "=== engine tests"; timeout 600 python -m pytest tests/test_gpu_engine.py -q -x 2>&1 | tail -15
echo "=== speed"; timeout 300 python tools/tc_speed.py tf32 2>&1 | tail -20
echo "=== gpu tests"; timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
echo "=== cfg5 n=262144 tf32"; timeout 600 python bench.py --n 262144 --steps 5 --warmup 3 --mode tf32 --alt-mode tf32x3 --no-e2e --no-cpu 2>&1 | tail -1
echo "=== cfg3 n=262144 tf32"; timeout 600 python bench.py --workload cfg3 --n 262144 --steps 5 --warmup 3 --mode tf32 --alt-mode tf32x3 --no-e2e --no-cpu 2>&1 | tail -1
} > gpurun_out/run10.log 2>&1
