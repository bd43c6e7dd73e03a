#!/bin/bash
mkdir -p gpurun_out
{
echo "=== fused tests"; timeout 600 python -m pytest tests/test_gpu_fused.py -m gpu -q 2>&1 | tail -15
echo "=== all gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], "it/s", d["ms_per_step"], "ms", d["roofline"]["phase_ms_per_step"], "launches", d["gpu_launches"])'
B="python bench.py --workload cfg5 --n 524288 --k 64 --steps 5 --warmup 3 --mode tf32 --alt-mode= --no-e2e --no-cpu"
echo "=== fit n=524288 f=8192 k=64 fused"; timeout 600 $B 2>&1 | tail -1 | (python -c "$P" || true)
echo "=== fit n=524288 f=8192 k=64 unfused"; KLNMF_FUSED=0 timeout 600 $B 2>&1 | tail -1 | (python -c "$P" || true)
} > gpurun_out/run23.log 2>&1
cat gpurun_out/run23.log | cut -c1-400
