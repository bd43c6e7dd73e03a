#!/bin/bash
mkdir -p gpurun_out
{ echo "=== property tests"; timeout 900 python -m pytest tests/test_gpu_properties.py -m gpu -q -x --durations=12 2>&1 | tail -40; } > gpurun_out/run45.log 2>&1
cut -c1-300 gpurun_out/run45.log
