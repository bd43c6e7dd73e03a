#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --n 65536 --steps 2 --warmup 1 --mode tf32 --alt-mode= --no-e2e --no-cpu"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_cfg5_tf32.csv $B > gpurun_out/ncu_list.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 1 -c 3 -f -o gpurun_out/prof_cfg5_tf32 $B > gpurun_out/ncu_full.log 2>&1
tail -n 3 gpurun_out/ncu_list.log gpurun_out/ncu_full.log
