#!/bin/bash
mkdir -p gpurun_out
{
echo "=== fused tests"; timeout 300 python -m pytest tests/test_gpu_fused.py -m gpu -q -x 2>&1 | tail -3
for d in 0 15; do
echo "=== dbg=$d n=524288"; KLNMF_F256_DBG=$d timeout 300 python bench.py --workload cfg3 --n 524288 --no-cpu --no-e2e --alt-mode= 2>&1 | tail -1 | grep -o '"ms_per_step": [0-9.]*'
done
echo "=== cfg3 full"; timeout 300 python bench.py --workload cfg3 --no-cpu --no-e2e --alt-mode= 2>&1 | tail -1 | grep -o '"ms_per_step": [0-9.]*'
echo "=== cfg3 k=192"; timeout 300 python bench.py --workload cfg3 --k 192 --no-cpu --no-e2e --alt-mode= 2>&1 | tail -1 | grep -o '"ms_per_step": [0-9.]*'
} > gpurun_out/run36.log 2>&1
cat gpurun_out/run36.log
