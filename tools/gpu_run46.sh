#!/bin/bash
mkdir -p gpurun_out
{ timeout 900 python tools/accuracy_vs_shape.py 2>&1 | tail -20; } > gpurun_out/run46.log 2>&1
cat gpurun_out/run46.log
