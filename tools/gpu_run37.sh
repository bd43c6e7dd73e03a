#!/bin/bash
mkdir -p gpurun_out
{
echo "=== fused tests"; timeout 300 python -m pytest tests/test_gpu_fused.py -m gpu -q -x 2>&1 | tail -3
for d in 0 16 15 31; do
echo "=== dbg=$d n=524288"; KLNMF_F256_DBG=$d timeout 300 python bench.py --workload cfg3 --n 524288 --no-cpu --no-e2e --alt-mode= 2>&1 | tail -1 | grep -o '"ms_per_step": [0-9.]*'
done
for k in 160 224; do
echo "=== cfg3 k=$k fused"; timeout 300 python bench.py --workload cfg3 --k $k --no-cpu --no-e2e --alt-mode= 2>&1 | tail -1 | grep -o '"ms_per_step": [0-9.]*'
echo "=== cfg3 k=$k unfused"; KLNMF_FUSED256=0 timeout 300 python bench.py --workload cfg3 --k $k --no-cpu --no-e2e --alt-mode= 2>&1 | tail -1 | grep -o '"ms_per_step": [0-9.]*'
done
} > gpurun_out/run37.log 2>&1
cat gpurun_out/run37.log
