#!/bin/bash
mkdir -p gpurun_out
{
echo "=== k512 ragged parity"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k k512 2>&1 | tail -5
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["roofline"]["phase_ms_per_step"], d["roofline"]["frac"], d["roofline"].get("issued_frac"))'
echo "=== cfg5 n=262144 tf32x3 phases"; KLNMF_PROFILE=1 timeout 300 python bench.py --n 262144 --mode tf32x3 --no-cpu --no-e2e --alt-mode= 2>&1 | tail -1 | python -c "$P"
echo "=== cfg3 tf32x3 phases"; KLNMF_PROFILE=1 timeout 300 python bench.py --workload cfg3 --mode tf32x3 --no-cpu --no-e2e --alt-mode= 2>&1 | tail -1 | python -c "$P"
echo "=== cfg5 n=262144 fp64 phases"; KLNMF_PROFILE=1 timeout 300 python bench.py --n 65536 --mode fp64 --no-cpu --no-e2e --alt-mode= --steps 2 --warmup 1 2>&1 | tail -1 | python -c "$P"
} > gpurun_out/run48.log 2>&1
cut -c1-500 gpurun_out/run48.log
