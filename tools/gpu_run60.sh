#!/bin/bash
mkdir -p gpurun_out
{
echo "=== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
echo "=== bench cfg5 n=1M e2e via learner"; timeout 900 python bench.py --n 1000000 --steps 5 --warmup 3 --no-cpu --alt-mode= 2>&1 | tail -1 | grep -o '"e2e": {[^}]*}' | cut -c1-700
echo "=== bench cfg3 e2e"; timeout 900 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu --alt-mode= 2>&1 | tail -1 | grep -o '"e2e": {[^}]*}' | cut -c1-300
} > gpurun_out/run60.log 2>&1
cat gpurun_out/run60.log
