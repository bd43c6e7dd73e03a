#!/bin/bash
# the driver's round-end commands at full size: default bench (cfg5, n=4M on one GPU) + reference arm
mkdir -p gpurun_out
{
echo "=== default bench"; S=$(date +%s); timeout 1500 python bench.py 2>&1 | tail -25; echo "wall $(( $(date +%s) - S )) s"
echo "=== reference arm"; S=$(date +%s); timeout 900 python bench.py --impl reference 2>&1 | tail -3; echo "wall $(( $(date +%s) - S )) s"
} > gpurun_out/run7.log 2>&1
exit 0
