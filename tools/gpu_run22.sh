#!/bin/bash
mkdir -p gpurun_out
{
echo "=== fused tests"; timeout 600 python -m pytest tests/test_gpu_fused.py -m gpu -q 2>&1 | tail -3
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], "it/s", d["ms_per_step"], "ms", "launches", d["gpu_launches"])'
B="python bench.py --workload cfg3 --n 524288 --k 128 --steps 5 --warmup 3 --mode tf32 --alt-mode= --no-e2e --no-cpu"
for V in "KLNMF_FUSED_V=0 KLNMF_FUSED_LA=1" "KLNMF_FUSED_V=0 KLNMF_FUSED_LA=2" "KLNMF_FUSED_V=1 KLNMF_FUSED_LA=1" "KLNMF_FUSED_V=1 KLNMF_FUSED_LA=2"; do
echo "=== k=128 $V"; env $V timeout 600 $B 2>&1 | tail -1 | (python -c "$P" || true)
done
B="python bench.py --workload cfg3 --n 524288 --k 64 --steps 5 --warmup 3 --mode tf32 --alt-mode= --no-e2e --no-cpu"
for V in "KLNMF_FUSED_LA=1" "KLNMF_FUSED_LA=2"; do
echo "=== k=64 $V"; env $V timeout 600 $B 2>&1 | tail -1 | (python -c "$P" || true)
done
} > gpurun_out/run22.log 2>&1
cat gpurun_out/run22.log | cut -c1-300
