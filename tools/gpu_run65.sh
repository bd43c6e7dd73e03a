#!/bin/bash
mkdir -p gpurun_out
{ timeout 900 python -m pytest tests -m gpu -q -k tf32r 2>&1 | grep -E "^FAILED|^E  .*(assert|Mismatch|Max rel|rel_fro)|passed|failed" | head -60; } > gpurun_out/run65.log 2>&1
cut -c1-260 gpurun_out/run65.log
