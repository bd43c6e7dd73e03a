#!/bin/bash
mkdir -p gpurun_out
{
echo "=== tests"; timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -4
timeout 900 python tools/e2e_breakdown.py 2>&1 | tail -3
} > gpurun_out/run31.log 2>&1
cat gpurun_out/run31.log | cut -c1-600
