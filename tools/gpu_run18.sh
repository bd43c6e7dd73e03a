#!/bin/bash
mkdir -p gpurun_out
{
echo "=== fused tests"; timeout 600 python -m pytest tests/test_gpu_fused.py -m gpu -q -x 2>&1 | tail -30
echo "=== all gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
} > gpurun_out/run18.log 2>&1
tail -60 gpurun_out/run18.log | cut -c1-400
