#!/bin/bash
# where does the 2x DRAM read volume come from?  dram bytes per kernel under L2-promotion / prefetch variants
mkdir -p gpurun_out
B="python bench.py --n 131072 --steps 1 --warmup 1 --mode tf32 --alt-mode= --no-e2e --no-cpu"
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum"
{
for V in "KLNMF_X=0" "KLNMF_TC_NOPF=1" "KLNMF_TC_PROMO=2" "KLNMF_TC_PROMO=0" "KLNMF_TC_PROMO=2 KLNMF_TC_NOPF=1"; do
echo "=== $V"
env $V timeout 600 ncu --metrics $M --clock-control none -k regex:tc_gemm -s 3 -c 3 --csv $B 2>&1 | grep -v "^==" | python -c "
import sys,csv
rows=list(csv.reader(sys.stdin))
for r in rows:
    if len(r)>14 and r[0].isdigit(): print(r[4][30:75], r[12], r[14], r[13])
"
done
} > gpurun_out/run16.log 2>&1
cat gpurun_out/run16.log | cut -c1-300
