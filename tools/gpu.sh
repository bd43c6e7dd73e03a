#!/bin/bash
# One parametrised runner for the GPU box (replaces the per-session gpu_run*.sh scripts of round 1):
#
#   gpurun --timeout 900 -- 'tools/gpu.sh LABEL STEP [STEP ...]'
#
# Every STEP is one of (arguments separated by ':'), all output lands in gpurun_out/LABEL.log:
#   tests[:PYTEST_K]                  python -m pytest tests -m gpu -x -q [-k PYTEST_K]   (testsall: without -x)
#   smoke                             __graft_entry__.smoke()
#   phases:WORKLOAD:N:MODE[:ENV=V,..] bench.py --workload WORKLOAD --n N --mode MODE, prints ms/step + phase times
#   bench[:ARGS]                      python bench.py ARGS (comma separated), prints the JSON line
#   launches:NAME[:ARGS]              ncu launch list (gpu__time_duration) of bench.py ARGS -> gpurun_out/NAME.csv
#   ncu:KERNEL_REGEX:NAME[:ARGS]      ncu --set full of the 2nd launch matching KERNEL_REGEX -> gpurun_out/NAME.ncu-rep
#   py:SCRIPT[:ARGS]                  python SCRIPT ARGS
mkdir -p gpurun_out
LABEL=$1; shift
LOG=gpurun_out/$LABEL.log
P='import sys,json
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1])
    r=d.get("roofline") or {}
    print("ms/step %.3f  it/s %.3f  phases %s  frac %s  clocks %s" % (d["ms_per_step"], d["value"], {k: round(v, 3) for k, v in (r.get("phase_ms_per_step") or {}).items()}, r.get("frac"), d.get("clocks")))
except Exception as e:
    print("no bench line:", e)'
{
for STEP in "$@"; do
  IFS=':' read -r KIND A1 A2 A3 A4 <<< "$STEP"
  echo "=== $STEP"
  case $KIND in
    tests)   timeout 1500 python -m pytest tests -m gpu -x -q ${A1:+-k "$A1"} 2>&1 | tail -15 ;;
    testsall) timeout 1500 python -m pytest tests -m gpu -q ${A1:+-k "$A1"} 2>&1 | tail -40 ;;
    smoke)   timeout 600 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -12 ;;
    phases)  env KLNMF_PROFILE=1 $(echo "$A4" | tr ',' ' ') timeout 900 python bench.py --workload "$A1" ${A2:+--n "$A2"} --mode "$A3" --no-cpu --no-e2e --alt-mode= --no-extra 2>&1 | tail -1 | python -c "$P" ;;
    bench)   timeout 1800 python bench.py $(echo "$A1" | tr ',' ' ') 2>&1 | tail -3 ;;
    launches) timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/$A1.csv python bench.py $(echo "$A2" | tr ',' ' ') 2>&1 | tail -1 | cut -c1-200 ;;
    ncu)     timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:$A1" -s 1 -c 1 -f -o gpurun_out/$A2 python bench.py $(echo "$A3" | tr ',' ' ') 2>&1 | tail -2 | cut -c1-200 ;;
    py)      timeout 1500 python $A1 $(echo "$A2" | tr ',' ' ') 2>&1 | tail -40 ;;
    *)       echo "unknown step $KIND" ;;
  esac
done
} > "$LOG" 2>&1
cut -c1-600 "$LOG"
