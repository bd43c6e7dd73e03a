#!/bin/bash
# ratio contraction with the ratio written back in place (KLNMF_TC_QIP=1): parity, then time per phase
mkdir -p gpurun_out
{
echo "=== parity tests with QIP"; KLNMF_TC_QIP=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_engine.py tests/test_gpu_reference_api.py -m gpu -q -x 2>&1 | tail -3
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["roofline"]["phase_ms_per_step"])'
echo "=== cfg5 n=262144 base"; KLNMF_PROFILE=1 timeout 300 python bench.py --n 262144 --no-cpu --no-e2e --alt-mode= 2>&1 | tail -1 | python -c "$P"
echo "=== cfg5 n=262144 QIP XB=4"; KLNMF_TC_QIP=1 KLNMF_TC_XB=4 KLNMF_PROFILE=1 timeout 300 python bench.py --n 262144 --no-cpu --no-e2e --alt-mode= 2>&1 | tail -1 | python -c "$P"
echo "=== cfg5 n=262144 QIP XB=6"; KLNMF_TC_QIP=1 KLNMF_TC_XB=6 KLNMF_PROFILE=1 timeout 300 python bench.py --n 262144 --no-cpu --no-e2e --alt-mode= 2>&1 | tail -1 | python -c "$P"
echo "=== cfg3 unfused base"; KLNMF_FUSED256=0 KLNMF_PROFILE=1 timeout 300 python bench.py --workload cfg3 --no-cpu --no-e2e --alt-mode= 2>&1 | tail -1 | python -c "$P"
echo "=== cfg3 unfused QIP XB=4"; KLNMF_FUSED256=0 KLNMF_TC_QIP=1 KLNMF_TC_XB=4 KLNMF_PROFILE=1 timeout 300 python bench.py --workload cfg3 --no-cpu --no-e2e --alt-mode= 2>&1 | tail -1 | python -c "$P"
echo "=== cfg3 unfused QIP XB=6"; KLNMF_FUSED256=0 KLNMF_TC_QIP=1 KLNMF_TC_XB=6 KLNMF_PROFILE=1 timeout 300 python bench.py --workload cfg3 --no-cpu --no-e2e --alt-mode= 2>&1 | tail -1 | python -c "$P"
} > gpurun_out/run42.log 2>&1
cat gpurun_out/run42.log
