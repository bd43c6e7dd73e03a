#!/bin/bash
# knock-out timing of the ratio contraction (cfg5 shape, n=262144): which part paces a tile
mkdir -p gpurun_out
{
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["roofline"]["phase_ms_per_step"]["ratio"])'
for q in 0 1; do for d in 0 1 2 3 4 5 6 7; do
echo "=== QIP=$q dbg=$d"; KLNMF_TC_QIP=$q KLNMF_TC_XB=$((2+2*q)) KLNMF_TC_DBG=$d KLNMF_PROFILE=1 timeout 300 python bench.py --n 262144 --no-cpu --no-e2e --alt-mode= --steps 3 2>&1 | tail -1 | python -c "$P"
done; done
} > gpurun_out/run43.log 2>&1
cat gpurun_out/run43.log
