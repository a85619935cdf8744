#!/usr/bin/env bash
# quick regression check on a GPU box: mean-shift + heads tests, then the full bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_meanshift_gpu.py tests/test_heads_gpu.py tests/test_full_size_gpu.py -q -m gpu --timeout 600 2>&1 | tail -5
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_check.json 2> gpurun_out/bench_check.err
tail -3 gpurun_out/bench_check.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_check.json"))
print(round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1))
print({k: (round(v["value"],1), round(v["ms_per_step"],2)) for k,v in d["configs"].items()})
print(d.get("densefusion_heads")); print(d["stock_gpu_baseline"].get("value"), d["cpu_baseline"]["value"])
PY
