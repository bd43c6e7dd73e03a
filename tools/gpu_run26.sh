#!/bin/bash
mkdir -p gpurun_out
{
echo "=== evaluation tests"; timeout 600 python -m pytest tests/test_gpu_evaluation.py -m gpu -q 2>&1 | tail -15
} > gpurun_out/run26.log 2>&1
cat gpurun_out/run26.log | cut -c1-400
