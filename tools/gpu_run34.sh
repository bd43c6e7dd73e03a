#!/bin/bash
# k=256 cluster kernel with the Q exchange by cp.async.bulk through distributed shared memory
mkdir -p gpurun_out
{
echo "=== fused tests"; timeout 300 python -m pytest tests/test_gpu_fused.py -m gpu -q -x 2>&1 | tail -8
echo "=== cfg3 fused256 on";  timeout 300 python bench.py --workload cfg3 --no-cpu --no-e2e --alt-mode= 2>&1 | tail -1 | cut -c1-2200
echo "=== cfg3 k=192 on"; timeout 300 python bench.py --workload cfg3 --k 192 --no-cpu --no-e2e --alt-mode= 2>&1 | tail -1 | cut -c 1-400
} > gpurun_out/run34.log 2>&1
cat gpurun_out/run34.log
