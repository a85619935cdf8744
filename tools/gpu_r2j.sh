#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 300 python tools/la_diag.py 2>&1 | tail -8 | tee gpurun_out/r2j_la_diag.log
for i in 1 2 3; do timeout 300 python bench.py --steps 20 --warmup 5 --quick > gpurun_out/r2j_bench_$i.json 2> gpurun_out/r2j_bench_$i.err; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2j_bench_*.json")):
    d=json.load(open(f)); print(f, round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1))
PY
