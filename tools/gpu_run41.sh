#!/bin/bash
# ratio contraction: rolling L2 prefetch of the X chunks (KLNMF_TC_PFD = distance in chunks), time and DRAM bytes
mkdir -p gpurun_out
{
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["roofline"]["phase_ms_per_step"])'
for d in 0 4 8 16; do
echo "=== cfg5 n=262144 PFD=$d"; KLNMF_TC_PFD=$d KLNMF_PROFILE=1 timeout 300 python bench.py --n 262144 --no-cpu --no-e2e --alt-mode= 2>&1 | tail -1 | python -c "$P"
done
for d in 0 8; do
echo "=== cfg3 unfused PFD=$d"; KLNMF_FUSED256=0 KLNMF_TC_PFD=$d KLNMF_PROFILE=1 timeout 300 python bench.py --workload cfg3 --no-cpu --no-e2e --alt-mode= 2>&1 | tail -1 | python -c "$P"
done
for d in 0 8; do
echo "=== dram bytes PFD=$d"
KLNMF_TC_PFD=$d timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:tc_gemm -s 3 -c 3 python bench.py --n 262144 --steps 2 --warmup 1 --no-cpu --no-e2e --alt-mode= 2>&1 | grep -E "tc_gemm_kernel|dram__|gpu__time" | cut -c1-120
done
} > gpurun_out/run41.log 2>&1
cat gpurun_out/run41.log
