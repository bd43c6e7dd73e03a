#!/bin/bash
# ratio epilogue with TMA-staged X/Q: correctness + speed vs the old epilogue
mkdir -p gpurun_out
{
echo "=== gpu tests"; timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
echo "=== bench cfg5 n=262144 tf32 (XT)"; timeout 600 python bench.py --n 262144 --steps 5 --warmup 3 --mode tf32 --alt-mode '' --no-e2e --no-cpu 2>&1 | tail -2
echo "=== bench cfg5 n=262144 tf32 (old epilogue)"; KLNMF_TC_NO_XT=1 timeout 600 python bench.py --n 262144 --steps 5 --warmup 3 --mode tf32 --alt-mode '' --no-e2e --no-cpu 2>&1 | tail -2
echo "=== bench cfg3 n=262144 tf32 (XT)"; timeout 600 python bench.py --workload cfg3 --n 262144 --steps 5 --warmup 3 --mode tf32 --alt-mode '' --no-e2e --no-cpu 2>&1 | tail -2
} > gpurun_out/run5.log 2>&1
tail -30 gpurun_out/run5.log
