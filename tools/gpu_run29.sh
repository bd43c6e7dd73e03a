#!/bin/bash
mkdir -p gpurun_out
{
echo "=== cfg2 test"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k cfg2 2>&1 | tail -5
echo "=== small configs"; timeout 900 python tools/small_configs.py 2>&1 | tail -12
} > gpurun_out/run29.log 2>&1
cat gpurun_out/run29.log | cut -c1-300
