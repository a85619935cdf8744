#!/usr/bin/env bash
# final validation of the round: the whole GPU suite, sanitizer, full bench line, reference arm
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu --timeout 900 2>&1 | tail -6 > gpurun_out/final_tests.log
cat gpurun_out/final_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
bash tools/sanitizer_r02.sh 2>&1 | grep "^=="
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_full.json 2> gpurun_out/bench_r02_full.err
tail -3 gpurun_out/bench_r02_full.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02_refarm.json 2> gpurun_out/bench_r02_refarm.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r02_full.json"))
print(round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "hbm_frac", round(d["frames_per_s_hbm_frac"]["value"],3), "roofline", round(d["roofline"]["frac"],3))
print({k: (round(v["value"],1), round(v["ms_per_step"],2)) for k,v in d["configs"].items()})
print(d["meanshift_modes"]); print(d["cpu_baseline"]["value"], d["stock_gpu_baseline"].get("value"), d["configs"]["ycb_b16"].get("cpu_meanshift"))
print(d["stage_ms_per_batch"]); print([ (r["kernel"][:30], round(r.get("ms_per_batch",0),3)) for r in d["rooflines"]])
r=json.load(open("gpurun_out/bench_r02_refarm.json")); print("ref arm", r["value"])
PY
