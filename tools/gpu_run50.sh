#!/bin/bash
mkdir -p gpurun_out
{
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["roofline"]["phase_ms_per_step"])'
for fm in 0 1; do
echo "=== FASTMATH=$fm cfg5 n=262144 tf32x3"; KLNMF_TC_FASTMATH=$fm KLNMF_PROFILE=1 timeout 300 python bench.py --n 262144 --mode tf32x3 --no-cpu --no-e2e --alt-mode= 2>&1 | tail -1 | python -c "$P"
echo "=== FASTMATH=$fm cfg3 tf32x3"; KLNMF_TC_FASTMATH=$fm KLNMF_PROFILE=1 timeout 300 python bench.py --workload cfg3 --mode tf32x3 --no-cpu --no-e2e --alt-mode= 2>&1 | tail -1 | python -c "$P"
echo "=== FASTMATH=$fm accuracy"; KLNMF_TC_FASTMATH=$fm timeout 600 python tools/accuracy_vs_shape.py 2>&1 | grep tf32x3
done
} > gpurun_out/run50.log 2>&1
cut -c1-400 gpurun_out/run50.log
