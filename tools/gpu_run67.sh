#!/bin/bash
mkdir -p gpurun_out
{ timeout 300 python tools/debug_tf32r.py 2>&1 | tail -12; echo "--- center off"; KLNMF_CENTER=0 timeout 300 python tools/debug_tf32r.py 2>&1 | grep tf32r; } > gpurun_out/run67.log 2>&1
cat gpurun_out/run67.log
