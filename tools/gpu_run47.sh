#!/bin/bash
# centered ratio (Q - 1): full gpu tests, accuracy against the FP64 mode by shape, per-phase times
mkdir -p gpurun_out
{
echo "=== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
echo "=== accuracy centered"; timeout 600 python tools/accuracy_vs_shape.py 2>&1 | tail -14
echo "=== accuracy plain (KLNMF_CENTER=0)"; KLNMF_CENTER=0 timeout 600 python tools/accuracy_vs_shape.py 2>&1 | tail -14
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["roofline"]["phase_ms_per_step"], d["alt_modes"])'
echo "=== cfg5 n=262144"; KLNMF_PROFILE=1 timeout 300 python bench.py --n 262144 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "$P"
echo "=== cfg5 n=262144 plain"; KLNMF_CENTER=0 KLNMF_PROFILE=1 timeout 300 python bench.py --n 262144 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "$P"
echo "=== cfg3"; KLNMF_PROFILE=1 timeout 300 python bench.py --workload cfg3 --no-cpu --no-e2e --alt-mode= 2>&1 | tail -1 | python -c "$P"
} > gpurun_out/run47.log 2>&1
cut -c1-400 gpurun_out/run47.log
