#!/bin/bash
mkdir -p gpurun_out
{
echo "=== all gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], "it/s", d["ms_per_step"], "ms", d["roofline"]["phase_ms_per_step"], d["roofline"]["frac"])'
B="python bench.py --workload cfg4 --n 262144 --steps 5 --warmup 3 --mode tf32 --alt-mode= --no-e2e --no-cpu"
echo "=== cfg4 n=262144 blocked-CSC"; timeout 600 $B 2>&1 | tail -1 | (python -c "$P" || true)
echo "=== cfg4 n=262144 atomics"; KLNMF_SPARSE_ATOMICS=1 timeout 600 $B 2>&1 | tail -1 | (python -c "$P" || true)
S="python bench.py --workload cfg4 --n 131072 --steps 2 --warmup 1 --mode tf32 --alt-mode= --no-e2e --no-cpu"
echo "=== ncu full sparse"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:sparse_ -s 2 -c 2 -f -o gpurun_out/r1_full_cfg4_n131072_v3 $S 2>&1 | tail -2
} > gpurun_out/run28.log 2>&1
cat gpurun_out/run28.log | cut -c1-500
