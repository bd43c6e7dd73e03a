#!/bin/bash
mkdir -p gpurun_out
{
echo "=== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
echo "=== smoke"; timeout 600 python __graft_entry__.py smoke 2>&1 | tail -10
echo "=== cfg5 n=524288 bench with alt modes"; timeout 600 python bench.py --n 524288 --no-cpu --no-e2e 2>&1 | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["alt_modes"])'
} > gpurun_out/run64.log 2>&1
cut -c1-600 gpurun_out/run64.log
