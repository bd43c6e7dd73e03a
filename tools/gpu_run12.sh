#!/bin/bash
# session-3 baseline: tests, smoke, ncu launch list + full captures (dense panel shape and sparse), default bench
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,power.limit --format=csv
echo "=== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
echo "=== smoke"; timeout 600 python __graft_entry__.py smoke 2>&1 | tail -8
B="python bench.py --n 262144 --steps 2 --warmup 1 --mode tf32 --alt-mode= --no-e2e --no-cpu"
echo "=== ncu list dense"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r1_launches_cfg5_n262144_tf32.csv $B 2>&1 | tail -2
echo "=== ncu full dense"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 3 -c 3 -f -o gpurun_out/r1_full_cfg5_n262144_tf32 $B 2>&1 | tail -3
S="python bench.py --workload cfg4 --n 131072 --steps 2 --warmup 1 --mode tf32 --alt-mode= --no-e2e --no-cpu"
echo "=== ncu list sparse"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r1_launches_cfg4_n131072.csv $S 2>&1 | tail -2
echo "=== ncu full sparse"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:sparse_ -s 2 -c 2 -f -o gpurun_out/r1_full_cfg4_n131072 $S 2>&1 | tail -3
echo "=== bench cfg4 n=262144"; timeout 900 python bench.py --workload cfg4 --n 262144 --steps 3 --warmup 3 --alt-mode= 2>&1 | tail -1
echo "=== bench cfg3 full"; timeout 900 python bench.py --workload cfg3 --steps 5 --warmup 3 2>&1 | tail -1
echo "=== reference arm"; timeout 900 python bench.py --impl reference --steps 5 --warmup 3 2>&1 | tail -1
echo "=== default bench"; timeout 1500 python bench.py 2>&1 | tail -1
} > gpurun_out/run12.log 2>&1
tail -60 gpurun_out/run12.log
