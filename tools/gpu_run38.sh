#!/bin/bash
mkdir -p gpurun_out
{
echo "=== fused tests"; timeout 300 python -m pytest tests/test_gpu_fused.py -m gpu -q -x 2>&1 | tail -3
for d in 0 32 15; do
echo "=== dbg=$d n=524288"; KLNMF_F256_DBG=$d timeout 300 python bench.py --workload cfg3 --n 524288 --no-cpu --no-e2e --alt-mode= 2>&1 | tail -1 | grep -o '"ms_per_step": [0-9.]*'
done
echo "=== dram bytes"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:fused_coef256 -s 1 -c 1 python bench.py --workload cfg3 --n 262144 --steps 2 --warmup 1 --no-cpu --no-e2e --alt-mode= 2>&1 | grep -E "dram__|gpu__time"
} > gpurun_out/run38.log 2>&1
cat gpurun_out/run38.log
