#!/bin/bash
mkdir -p gpurun_out
{
echo "=== parity tests, default"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -25
echo "=== parity tests, KLNMF_TC_FASTMATH=0"; KLNMF_TC_FASTMATH=0 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -5
} > gpurun_out/run52.log 2>&1
cut -c1-600 gpurun_out/run52.log
