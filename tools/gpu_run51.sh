#!/bin/bash
mkdir -p gpurun_out
{
echo "=== gpu tests (graph replay on)"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
echo "=== small configs, graphs on"; timeout 600 python tools/small_configs.py 2>&1 | tail -12
echo "=== small configs, KLNMF_GRAPH=0"; KLNMF_GRAPH=0 timeout 600 python tools/small_configs.py 2>&1 | tail -12
echo "=== cfg3"; timeout 300 python bench.py --workload cfg3 --no-cpu --no-e2e --alt-mode= --steps 20 2>&1 | tail -1 | grep -o '"ms_per_step": [0-9.]*\|"gpu_launches": [0-9]*'
} > gpurun_out/run51.log 2>&1
cut -c1-300 gpurun_out/run51.log
