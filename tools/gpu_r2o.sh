#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_full.json 2> gpurun_out/bench_r02_full.err
tail -3 gpurun_out/bench_r02_full.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02_refarm.json 2> gpurun_out/bench_r02_refarm.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r02_full.json"))
print(round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1))
print({k: (round(v["value"],1), round(v["ms_per_step"],2)) for k,v in d["configs"].items()})
print(d["meanshift_modes"]); print(d["cpu_baseline"]["value"], d["stock_gpu_baseline"].get("value"))
r=json.load(open("gpurun_out/bench_r02_refarm.json")); print("ref arm", r["value"], r["cpu_baseline"]["detail"])
PY
