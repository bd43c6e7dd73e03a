#!/bin/bash
mkdir -p gpurun_out
{ timeout 300 python tools/debug_blocks.py 2>&1 | tail -4
  timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "dense_stack or learner" 2>&1 | tail -5; } > gpurun_out/run59.log 2>&1
cat gpurun_out/run59.log
