#!/bin/bash
mkdir -p gpurun_out
{
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], "it/s", d["ms_per_step"], "ms", d["roofline"]["phase_ms_per_step"], "launches", d["gpu_launches"])'
for K in 128 64 50; do
B="python bench.py --workload cfg3 --n 524288 --k $K --steps 5 --warmup 3 --mode tf32 --alt-mode= --no-e2e --no-cpu"
echo "=== transform n=524288 f=4096 k=$K fused";   timeout 600 $B 2>&1 | tail -1 | (python -c "$P" || true)
echo "=== transform n=524288 f=4096 k=$K unfused"; KLNMF_FUSED=0 timeout 600 $B 2>&1 | tail -1 | (python -c "$P" || true)
done
B2="python bench.py --workload cfg3 --n 524288 --k 128 --steps 2 --warmup 1 --mode tf32 --alt-mode= --no-e2e --no-cpu"
echo "=== ncu full fused"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:fused_coef -s 1 -c 1 -f -o gpurun_out/r1_full_fused_k128 $B2 2>&1 | tail -2
} > gpurun_out/run19.log 2>&1
cat gpurun_out/run19.log | cut -c1-400
