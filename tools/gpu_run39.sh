#!/bin/bash
mkdir -p gpurun_out
{
echo "=== fused tests"; timeout 300 python -m pytest tests/test_gpu_fused.py -m gpu -q -x 2>&1 | tail -3
echo "=== n=524288"; timeout 300 python bench.py --workload cfg3 --n 524288 --no-cpu --no-e2e --alt-mode= 2>&1 | tail -1 | grep -o '"ms_per_step": [0-9.]*'
echo "=== cfg3 full"; timeout 300 python bench.py --workload cfg3 --no-cpu --no-e2e --alt-mode= 2>&1 | tail -1 | grep -o '"ms_per_step": [0-9.]*'
echo "=== cfg3 k=192"; timeout 300 python bench.py --workload cfg3 --k 192 --no-cpu --no-e2e --alt-mode= 2>&1 | tail -1 | grep -o '"ms_per_step": [0-9.]*'
} > gpurun_out/run39.log 2>&1
cat gpurun_out/run39.log
