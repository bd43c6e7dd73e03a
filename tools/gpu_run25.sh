#!/bin/bash
# round-1 record run: tests, smoke, launch list + full capture at the cfg5 panel shape, default bench, cfg3 / cfg4 benches
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,power.limit --format=csv
echo "=== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
echo "=== smoke"; timeout 600 python __graft_entry__.py smoke 2>&1 | tail -8
B="python bench.py --n 262144 --steps 2 --warmup 1 --mode tf32 --alt-mode= --no-e2e --no-cpu"
echo "=== ncu list dense"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r1_launches_cfg5_n262144_tf32_v2.csv $B 2>&1 | tail -1 | cut -c1-200
echo "=== ncu full dense"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 3 -c 3 -f -o gpurun_out/r1_full_cfg5_n262144_tf32_v2 $B 2>&1 | tail -1
echo "=== default bench"; timeout 1500 python bench.py 2>&1 | tail -1
echo "=== bench cfg3 full"; timeout 900 python bench.py --workload cfg3 --steps 5 --warmup 3 2>&1 | tail -1
echo "=== bench cfg4 n=500000"; timeout 900 python bench.py --workload cfg4 --n 500000 --steps 5 --warmup 3 --alt-mode= 2>&1 | tail -1
echo "=== reference arm"; timeout 900 python bench.py --impl reference --steps 5 --warmup 3 2>&1 | tail -1
} > gpurun_out/run25.log 2>&1
tail -30 gpurun_out/run25.log | cut -c1-3000
