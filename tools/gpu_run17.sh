#!/bin/bash
mkdir -p gpurun_out
{
echo "=== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["roofline"]["phase_ms_per_step"], d.get("alt_modes"))'
for XB in 4 2; do
B="python bench.py --n 262144 --steps 5 --warmup 3 --mode tf32 --alt-mode= --no-e2e --no-cpu"
echo "=== cfg5 XB=$XB"; KLNMF_TC_XB=$XB timeout 600 $B 2>&1 | tail -1 | (python -c "$P" || true)
echo "=== cfg3 XB=$XB"; KLNMF_TC_XB=$XB timeout 600 $B --workload cfg3 2>&1 | tail -1 | (python -c "$P" || true)
done
B2="python bench.py --n 262144 --steps 2 --warmup 1 --mode tf32 --alt-mode= --no-e2e --no-cpu"
echo "=== ncu full ratio"
timeout 1500 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:Lb0ELb1ELb0ELb1ELi2 -s 1 -c 1 -f -o gpurun_out/r1_full_ratio_v5 $B2 2>&1 | tail -2
} > gpurun_out/run17.log 2>&1
tail -40 gpurun_out/run17.log | cut -c1-600
