#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python tools/la_diag.py 2>&1 | tail -12 | tee gpurun_out/r2n_la_diag.log
for i in 1 2; do timeout 300 python bench.py --steps 20 --warmup 5 --quick > gpurun_out/r2n_bench_$i.json 2> gpurun_out/r2n_bench_$i.err; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2n_bench_*.json")):
    d=json.load(open(f)); print(f, round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1))
PY
for i in 1 2 3; do timeout 300 /usr/local/cuda/bin/compute-sanitizer --tool racecheck python -m pytest tests/test_mlp_gpu.py -q -m gpu -x -k "test_fp_first" 2>&1 | grep -E "passed|failed|assert tensor|RACECHECK SUMMARY" | tr '\n' ' '; echo; done
