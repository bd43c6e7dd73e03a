#!/bin/bash
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/mc_bench tools/mc_bench.cu > gpurun_out/mc_bench.log 2>&1
timeout 25 gpurun_out/mc_bench 4000 48 >> gpurun_out/mc_bench.log 2>&1
echo "rc=$?" >> gpurun_out/mc_bench.log
rm -f gpurun_out/mc_bench
cat gpurun_out/mc_bench.log
