#!/bin/bash
# first GPU session: box facts, fp64 path parity, tcgen05 bring-up
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
nproc; free -g | head -2
python -c "import __graft_entry__ as g; print(g.build())"
echo "=== engine fp64"
timeout 600 python -m pytest tests/test_gpu_engine.py -q -k "fp64" -x 2>&1 | tail -8
echo "=== tc_debug tf32"
timeout 300 python tools/tc_debug.py tf32 2>&1 | tail -120
echo "=== engine tf32"
timeout 900 python -m pytest tests/test_gpu_engine.py -q -k "not fp64" 2>&1 | tail -40
echo "=== parity fp64"
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_api.py -q -k "fp64" 2>&1 | tail -40
echo "=== parity tf32 modes"
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_api.py -q -k "not fp64" 2>&1 | tail -40
} > gpurun_out/run1.log 2>&1
tail -150 gpurun_out/run1.log
