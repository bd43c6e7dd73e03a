#!/bin/bash
mkdir -p gpurun_out
{
B="python bench.py --n 262144 --steps 5 --warmup 3 --mode tf32 --alt-mode= --no-e2e --no-cpu"
echo "=== cfg5 relaxed"; timeout 600 $B 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['phase_ms_per_step'])"
echo "=== cfg5 release"; KLNMF_TC_RELAXED=0 timeout 600 $B 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['phase_ms_per_step'])"
echo "=== kswz32 detail"; KLNMF_TC_KSWZ32=1 timeout 600 python -m pytest "tests/test_gpu_engine.py::test_contract_is_exact_on_tf32_representable_inputs" -m gpu -q -x 2>&1 | grep -v "^$" | tail -25
B2="python bench.py --n 262144 --steps 2 --warmup 1 --mode tf32 --alt-mode= --no-e2e --no-cpu"
echo "=== ncu full ratio"
timeout 1500 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:Lb0ELb1ELb0ELb1ELi2 -s 1 -c 1 -f -o gpurun_out/r1_full_ratio_v2 $B2 2>&1 | tail -2
} > gpurun_out/run14.log 2>&1
tail -40 gpurun_out/run14.log | cut -c1-600
