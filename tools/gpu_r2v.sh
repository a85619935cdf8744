#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_heads_gpu.py tests/test_mlp_gpu.py -q -m gpu --timeout 600 -s 2>&1 | grep -E "passed|failed|fused mean|Error|error|assert" | tail -25
timeout 300 python bench.py --steps 20 --warmup 5 --quick > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2v_bench.json"))
r={x["kernel"][:12]:round(x.get("ms_per_batch",0),3) for x in d.get("rooflines",[])}
print(round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), r)
PY
