#!/bin/bash
mkdir -p gpurun_out
{ echo "=== default bench (tf32 + tf32r + tf32x3 at full cfg5)"; S=$(date +%s); timeout 900 python bench.py 2>&1 | tail -1; echo "wall $(( $(date +%s) - S )) s"; } > gpurun_out/run70.log 2>&1
python - <<'PY'
import json
for line in open('gpurun_out/run70.log'):
    if line.startswith('{"metric"'):
        d=json.loads(line)
        print(d["ms_per_step"], d["e2e"]["value"], d["e2e"]["seconds"], {k:(round(v["ms_per_step"],1), round(v["roofline_frac"],3)) for k,v in d["alt_modes"].items()}, d["clocks"])
PY
tail -1 gpurun_out/run70.log
