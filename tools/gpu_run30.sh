#!/bin/bash
mkdir -p gpurun_out
{ timeout 900 python tools/e2e_breakdown.py 2>&1 | tail -4; } > gpurun_out/run30.log 2>&1
cat gpurun_out/run30.log | cut -c1-600
