#!/bin/bash
mkdir -p gpurun_out
{
echo "=== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6
echo "=== smoke"; timeout 600 python __graft_entry__.py smoke 2>&1 | tail -10
} > gpurun_out/run66.log 2>&1
cut -c1-300 gpurun_out/run66.log
