#!/bin/bash
mkdir -p gpurun_out
{
echo "=== sweep"; timeout 300 python tools/tc_sweep.py 2>&1
echo "=== table"; timeout 600 python tools/tc_table.py tf32 tf32x3 2>&1
} > gpurun_out/run2.log 2>&1
tail -5 gpurun_out/run2.log
