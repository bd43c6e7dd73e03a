#!/bin/bash
mkdir -p gpurun_out
{
echo "=== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["roofline"]["phase_ms_per_step"])'
echo "=== cfg5 n=262144 tf32x3"; KLNMF_PROFILE=1 timeout 300 python bench.py --n 262144 --mode tf32x3 --no-cpu --no-e2e --alt-mode= 2>&1 | tail -1 | python -c "$P"
echo "=== cfg3 tf32x3"; KLNMF_PROFILE=1 timeout 300 python bench.py --workload cfg3 --mode tf32x3 --no-cpu --no-e2e --alt-mode= 2>&1 | tail -1 | python -c "$P"
echo "=== accuracy"; timeout 600 python tools/accuracy_vs_shape.py 2>&1 | grep tf32x3
} > gpurun_out/run56.log 2>&1
cut -c1-400 gpurun_out/run56.log
