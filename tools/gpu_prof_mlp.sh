#!/usr/bin/env bash
# ncu --set full of the 33 layer launches of one forward (tools/mlp_ab.py), then the full bench line
mkdir -p gpurun_out
TAG=${1:-r02b}
AB_LAUNCHES=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlp_layer_kernel -s 33 -c 33 -f -o gpurun_out/prof_mlp_${TAG} python tools/mlp_ab.py > gpurun_out/ncu_mlp_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_mlp_${TAG}.log
timeout 300 python tools/mlp_ab.py 2>&1 | grep -v "Warning\|_warn_once" | tail -3
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_check.json 2> gpurun_out/bench_check.err
tail -3 gpurun_out/bench_check.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_check.json"))
print(round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1))
print({k: (round(v["value"],1), round(v["ms_per_step"],2)) for k,v in d["configs"].items()})
print(d.get("densefusion_heads")); print(d["stock_gpu_baseline"].get("value"), d["cpu_baseline"]["value"])
print(d.get("stages_ms") or d.get("stages"))
PY
