#!/bin/bash
# lean MMA issuer + relaxed accumulator hand-back: tests, K-major/32B-atom experiment, benches
mkdir -p gpurun_out
{
echo "=== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
echo "=== engine tests with K-major 32B-atom swizzle"; KLNMF_TC_KSWZ32=1 timeout 600 python -m pytest tests/test_gpu_engine.py -m gpu -q 2>&1 | tail -8
echo "=== speed"; timeout 600 python tools/tc_speed.py 2>&1 | tail -18
echo "=== cfg5 n=262144 tf32"; timeout 600 python bench.py --n 262144 --steps 5 --warmup 3 --mode tf32 --alt-mode tf32x3 --no-e2e --no-cpu 2>&1 | tail -1
echo "=== cfg3 n=262144 tf32"; timeout 600 python bench.py --workload cfg3 --n 262144 --steps 5 --warmup 3 --mode tf32 --alt-mode tf32x3 --no-e2e --no-cpu 2>&1 | tail -1
B="python bench.py --n 262144 --steps 2 --warmup 1 --mode tf32 --alt-mode= --no-e2e --no-cpu"
echo "=== ncu full ratio"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 3 -c 1 -f -o gpurun_out/r1_full_ratio_v2 $B 2>&1 | tail -2
} > gpurun_out/run13.log 2>&1
tail -40 gpurun_out/run13.log | cut -c1-1500
