#!/bin/bash
# timing experiments on the k=256 cluster kernel: which resource paces a step
mkdir -p gpurun_out
{
for d in 0 1 2 4 8 3 6 15; do
echo "=== dbg=$d"; KLNMF_F256_DBG=$d timeout 300 python bench.py --workload cfg3 --n 524288 --no-cpu --no-e2e --alt-mode= 2>&1 | tail -1 | grep -o '"ms_per_step": [0-9.]*'
done
} > gpurun_out/run35.log 2>&1
cat gpurun_out/run35.log
