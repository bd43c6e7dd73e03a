#!/bin/bash
mkdir -p gpurun_out
{
echo "=== fused tests"; timeout 600 python -m pytest tests/test_gpu_fused.py -m gpu -q 2>&1 | tail -15
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], "it/s", d["ms_per_step"], "ms", "launches", d["gpu_launches"])'
for K in 128 64; do
B="python bench.py --workload cfg3 --n 524288 --k $K --steps 5 --warmup 3 --mode tf32 --alt-mode= --no-e2e --no-cpu"
echo "=== transform n=524288 f=4096 k=$K fused W-in-smem";   timeout 600 $B 2>&1 | tail -1 | (python -c "$P" || true)
echo "=== transform n=524288 f=4096 k=$K fused W-in-tmem";   KLNMF_FUSED_TS=1 timeout 600 $B 2>&1 | tail -1 | (python -c "$P" || true)
done
} > gpurun_out/run20.log 2>&1
cat gpurun_out/run20.log | cut -c1-300
