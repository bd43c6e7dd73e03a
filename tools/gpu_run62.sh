#!/bin/bash
mkdir -p gpurun_out
{
echo "=== fused tests"; timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_properties.py -m gpu -q -x 2>&1 | tail -4
for k in 128 64; do
echo "=== transform n=524288 f=4096 k=$k"; timeout 300 python bench.py --workload cfg3 --n 524288 --k $k --no-cpu --no-e2e --alt-mode= 2>&1 | tail -1 | grep -o '"ms_per_step": [0-9.]*'
done
echo "=== fit n=262144 f=8192 k=64"; timeout 300 python bench.py --n 262144 --k 64 --no-cpu --no-e2e --alt-mode= 2>&1 | tail -1 | grep -o '"ms_per_step": [0-9.]*'
} > gpurun_out/run62.log 2>&1
cat gpurun_out/run62.log
