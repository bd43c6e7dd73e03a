#!/bin/bash
mkdir -p gpurun_out
{
echo "=== engine"; timeout 900 python -m pytest tests/test_gpu_engine.py -q 2>&1 | tail -15
echo "=== parity"; timeout 1800 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_api.py -q 2>&1 | tail -60
echo "=== smoke"; timeout 600 python __graft_entry__.py smoke 2>&1 | tail -12
echo "=== bench small"; timeout 900 python bench.py --n 262144 --steps 3 --warmup 2 2>&1 | tail -5
} > gpurun_out/run3.log 2>&1
tail -5 gpurun_out/run3.log
