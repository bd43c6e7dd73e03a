#!/bin/bash
mkdir -p gpurun_out
{
echo "=== e2e breakdown n=1M cfg5 shape"; N=1000000 timeout 600 python tools/e2e_breakdown.py 2>&1 | tail -3
echo "=== bench cfg3 e2e"; timeout 900 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu --alt-mode= 2>&1 | tail -1 | grep -o '"e2e": {[^}]*}' | cut -c1-200
echo "=== bench cfg3 e2e again"; timeout 900 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu --alt-mode= 2>&1 | tail -1 | grep -o '"e2e": {[^}]*}' | cut -c1-200
nproc; free -g | head -2
} > gpurun_out/run54.log 2>&1
cat gpurun_out/run54.log
