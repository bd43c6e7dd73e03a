#!/bin/bash
mkdir -p gpurun_out
{
timeout 300 python tools/debug_tf32r.py 2>&1 | grep -A2 "^tf32r"
echo "=== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5
echo "=== smoke"; timeout 600 python __graft_entry__.py smoke 2>&1 | tail -3
} > gpurun_out/run68.log 2>&1
cut -c1-300 gpurun_out/run68.log
