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
#   ncu:KERNEL_REGEX:NAME:ARGS[:S,C]  ncu --set full of C launches matching KERNEL_REGEX after skipping S (default 1,1)
#                                     -> gpurun_out/NAME.ncu-rep
#   py:SCRIPT[:ARGS]                  python SCRIPT ARGS
#   dist:N[:ARGS[:ENV=V,..]]          bench.py ARGS under torchrun on N GPUs (+ NCCL transport summary)
#   distpy:N:SCRIPT[:ARGS]            any script under torchrun on N GPUs
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
PD='import sys,json
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1])
    r=d.get("roofline") or {}
    e=d.get("e2e") or {}
    print("N=%d ms/step %.3f it/s %.3f e2e %.3f it/s (%.2f s) e2e50 %s" % (d["n_gpus"], d["ms_per_step"], d["value"], e.get("value", 0), e.get("seconds", 0), (e.get("at_reference_iterations") or {}).get("value")))
    print("  phases min/max over ranks:", {k: [round(x, 3) for x in v] for k, v in (r.get("phase_ms_per_step_min_max_over_ranks") or {}).items()})
    print("  clocks", d.get("clocks"))
    for k, v in (d.get("workloads") or {}).items(): print(("  %s: %.3f ms/step %.2f it/s frac %s" % (k, v["ms_per_step"], v["value"], (v.get("roofline") or {}).get("frac"))) if "ms_per_step" in v else ("  %s: %s" % (k, {a: b for a, b in v.items() if a != "workload"})))
    for k, v in (d.get("alt_modes") or {}).items(): print("  alt %s: %.3f ms/step" % (k, v["ms_per_step"]))
except Exception as ex:
    print("no bench line:", ex)'
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
    ncu)     # ncu:KERNEL_REGEX:NAME:ARGS[:SKIP,COUNT]   (default: skip 1 matching launch, capture 1)
             SC=${A4:-1,1}; timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:$A1" -s ${SC%,*} -c ${SC#*,} -f -o gpurun_out/$A2 python bench.py $(echo "$A3" | tr ',' ' ') 2>&1 | tail -2 | cut -c1-200 ;;
    py)      timeout 1500 python $A1 $(echo "$A2" | tr ',' ' ') 2>&1 | tail -40 ;;
    dist)    # dist:N[:ARGS[:ENV=V,..]]  bench.py under torchrun on N GPUs; NCCL's transport choice is summarised from NCCL_DEBUG=INFO
             env NCCL_DEBUG=INFO NCCL_DEBUG_FILE=gpurun_out/$LABEL.nccl.%h.%p $(echo "$A3" | tr ',' ' ') timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$A1" --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus "$A1" $(echo "$A2" | tr ',' ' ') > gpurun_out/$LABEL.dist$A1.out 2>&1
             grep '^{' gpurun_out/$LABEL.dist$A1.out | tail -1 | python -c "$PD"
             cat gpurun_out/$LABEL.nccl.* 2>/dev/null | grep -o -E "via [A-Za-z0-9/_]+|NVLS[ a-zA-Z]*|Connected all (rings|trees)[^,]*|Using network [A-Za-z]+|[0-9]+ coll channels" | sort | uniq -c | sort -rn | head -12
             rm -f gpurun_out/$LABEL.nccl.* ;;
    distpy)  # distpy:N:SCRIPT[:ARGS]  any script under torchrun
             timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$A1" --master-addr 127.0.0.1 --master-port 29512 $A2 $(echo "$A3" | tr ',' ' ') 2>&1 | tail -30 ;;
    *)       echo "unknown step $KIND" ;;
  esac
done
} > "$LOG" 2>&1
cut -c1-600 "$LOG"
