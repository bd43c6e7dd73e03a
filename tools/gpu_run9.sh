#!/bin/bash
mkdir -p gpurun_out
{
echo "=== gpu tests"; timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
for pad in 32 0; do
echo "=== pad=$pad cfg5 n=262144 tf32"; KLNMF_LD_PAD=$pad timeout 600 python bench.py --n 262144 --steps 5 --warmup 3 --mode tf32 --alt-mode= --no-e2e --no-cpu 2>&1 | tail -1
echo "=== pad=$pad cfg3 n=262144 tf32"; KLNMF_LD_PAD=$pad timeout 600 python bench.py --workload cfg3 --n 262144 --steps 5 --warmup 3 --mode tf32 --alt-mode= --no-e2e --no-cpu 2>&1 | tail -1
done
} > gpurun_out/run9.log 2>&1
