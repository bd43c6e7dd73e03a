#!/bin/bash
mkdir -p gpurun_out
{
echo "=== accuracy"; timeout 600 python tools/accuracy_vs_shape.py 2>&1 | tail -20
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["roofline"]["phase_ms_per_step"])'
for m in tf32 tf32r; do
echo "=== cfg5 n=262144 $m"; KLNMF_PROFILE=1 timeout 300 python bench.py --n 262144 --mode $m --no-cpu --no-e2e --alt-mode= 2>&1 | tail -1 | python -c "$P"
done
echo "=== cfg3 tf32r"; KLNMF_PROFILE=1 timeout 300 python bench.py --workload cfg3 --mode tf32r --no-cpu --no-e2e --alt-mode= 2>&1 | tail -1 | python -c "$P"
echo "=== gpu tests (tf32 kernels touched)"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
} > gpurun_out/run63.log 2>&1
cut -c1-300 gpurun_out/run63.log
