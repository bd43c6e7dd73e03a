#!/bin/bash
mkdir -p gpurun_out
{
for i in 1 2; do
echo "=== bench cfg3 e2e"; timeout 900 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu --alt-mode= 2>&1 | tail -1 | grep -o '"e2e": {[^}]*}' | cut -c1-160
done
echo "=== bench cfg4 e2e"; timeout 900 python bench.py --workload cfg4 --n 500000 --steps 5 --warmup 3 --no-cpu --alt-mode= 2>&1 | tail -1 | grep -o '"e2e": {[^}]*}' | cut -c1-160
echo "=== bench cfg5 n=1M e2e"; timeout 900 python bench.py --n 1000000 --steps 5 --warmup 3 --no-cpu --alt-mode= 2>&1 | tail -1 | grep -o '"e2e": {[^}]*}' | cut -c1-160
} > gpurun_out/run55.log 2>&1
cat gpurun_out/run55.log
