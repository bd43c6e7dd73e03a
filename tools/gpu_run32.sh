#!/bin/bash
mkdir -p gpurun_out
{
echo "=== fused256 tests"; timeout 300 python -m pytest tests/test_gpu_fused.py -m gpu -q -x -k k256 2>&1 | tail -25
} > gpurun_out/run32.log 2>&1
cat gpurun_out/run32.log | cut -c1-400
