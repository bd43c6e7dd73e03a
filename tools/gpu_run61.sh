#!/bin/bash
# round-1 closing record run: tests, smoke, default bench (cfg5 full, tf32x3 alt, e2e through the learner), cfg3, cfg4, reference arm, small configs
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,power.limit --format=csv
echo "=== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
echo "=== smoke"; timeout 600 python __graft_entry__.py smoke 2>&1 | tail -8
echo "=== default bench"; S=$(date +%s); timeout 1500 python bench.py 2>&1 | tail -1; echo "wall $(( $(date +%s) - S )) s"
echo "=== bench cfg3 full"; timeout 900 python bench.py --workload cfg3 --steps 5 --warmup 3 2>&1 | tail -1
echo "=== bench cfg4 n=500000"; timeout 900 python bench.py --workload cfg4 --n 500000 --steps 5 --warmup 3 --alt-mode= 2>&1 | tail -1
echo "=== reference arm"; timeout 900 python bench.py --impl reference --steps 5 --warmup 3 2>&1 | tail -1
echo "=== small configs"; timeout 600 python tools/small_configs.py 2>&1 | tail -12
} > gpurun_out/run61.log 2>&1
tail -40 gpurun_out/run61.log | cut -c1-400
