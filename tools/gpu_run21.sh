#!/bin/bash
mkdir -p gpurun_out
{
echo "=== fused tests"; timeout 600 python -m pytest tests/test_gpu_fused.py -m gpu -q 2>&1 | tail -5
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], "it/s", d["ms_per_step"], "ms", "launches", d["gpu_launches"])'
for K in 128 64; do
B="python bench.py --workload cfg3 --n 524288 --k $K --steps 5 --warmup 3 --mode tf32 --alt-mode= --no-e2e --no-cpu"
echo "=== transform n=524288 f=4096 k=$K fused W-in-smem";   KLNMF_FUSED_TS=0 timeout 600 $B 2>&1 | tail -1 | (python -c "$P" || true)
echo "=== transform n=524288 f=4096 k=$K fused W-in-tmem";   timeout 600 $B 2>&1 | tail -1 | (python -c "$P" || true)
done
B2="python bench.py --workload cfg3 --n 524288 --k 64 --steps 2 --warmup 1 --mode tf32 --alt-mode= --no-e2e --no-cpu"
echo "=== ncu full fused k=64"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:fused_coef -s 1 -c 1 -f -o gpurun_out/r1_full_fused_k64_v2 $B2 2>&1 | tail -2
} > gpurun_out/run21.log 2>&1
cat gpurun_out/run21.log | cut -c1-300
