#!/bin/bash
# session-4 record run: tests, smoke, launch list + full captures of the final kernels, default bench, cfg3 / cfg4 benches
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,power.limit --format=csv
echo "=== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
echo "=== smoke"; timeout 600 python __graft_entry__.py smoke 2>&1 | tail -8
B="python bench.py --n 262144 --steps 2 --warmup 1 --mode tf32 --alt-mode= --no-e2e --no-cpu"
echo "=== ncu list dense"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r1_launches_cfg5_n262144_tf32_v3.csv $B 2>&1 | tail -1 | cut -c1-200
echo "=== ncu full dense"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 3 -c 3 -f -o gpurun_out/r1_full_cfg5_n262144_tf32_v3 $B 2>&1 | tail -1
echo "=== ncu full fused256"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_coef256 -s 1 -c 1 -f -o gpurun_out/r1_fused256_v5_cfg3_n262144 python bench.py --workload cfg3 --n 262144 --steps 2 --warmup 1 --no-cpu --no-e2e --alt-mode= 2>&1 | tail -1
echo "=== ncu full sparse"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sparse_ -s 2 -c 2 -f -o gpurun_out/r1_full_cfg4_n131072_v4 python bench.py --workload cfg4 --n 131072 --steps 2 --warmup 1 --alt-mode= --no-e2e --no-cpu 2>&1 | tail -1
echo "=== default bench"; S=$(date +%s); timeout 1500 python bench.py 2>&1 | tail -1; echo "wall $(( $(date +%s) - S )) s"
echo "=== bench cfg3 full"; timeout 900 python bench.py --workload cfg3 --steps 5 --warmup 3 2>&1 | tail -1
echo "=== bench cfg4 n=500000"; timeout 900 python bench.py --workload cfg4 --n 500000 --steps 5 --warmup 3 --alt-mode= 2>&1 | tail -1
echo "=== reference arm"; timeout 900 python bench.py --impl reference --steps 5 --warmup 3 2>&1 | tail -1
} > gpurun_out/run44.log 2>&1
tail -30 gpurun_out/run44.log | cut -c1-1200
