#!/bin/bash
mkdir -p gpurun_out
{
echo "=== dense stack on the device + parity file"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_api.py -m gpu -q -x 2>&1 | tail -8
} > gpurun_out/run58.log 2>&1
cut -c1-400 gpurun_out/run58.log
