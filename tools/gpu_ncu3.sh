#!/bin/bash
# ncu --set full of the k=256 cluster kernel (cfg3 shape, n=262144); label in $1
mkdir -p gpurun_out
L=${1:-r1_fused256_v2_cfg3_n262144}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_coef256 -s 1 -c 1 -f -o gpurun_out/$L python bench.py --workload cfg3 --n 262144 --steps 2 --warmup 1 --no-cpu --no-e2e --alt-mode= > gpurun_out/ncu3.log 2>&1
tail -n 2 gpurun_out/ncu3.log | cut -c1-300
