#!/bin/bash
mkdir -p gpurun_out
{ timeout 600 python -m pytest tests -m gpu -q -k "learner or stack or evaluation" 2>&1 | tail -4; } > gpurun_out/run69.log 2>&1
cat gpurun_out/run69.log
