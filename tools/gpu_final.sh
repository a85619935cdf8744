#!/usr/bin/env bash
# final validation of the round: the whole GPU suite, smoke, full bench line, reference arm, ncu launch list + layer
# kernel capture, sanitizer
mkdir -p gpurun_out
TAG=${1:-r02c}
timeout 1800 python -m pytest tests -q -m gpu --timeout 900 2>&1 | tail -6 > gpurun_out/final_tests.log
cat gpurun_out/final_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
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
print(d.get("densefusion_heads"))
r=json.load(open("gpurun_out/bench_r02_refarm.json")); print("ref arm", r["value"])
PY
B="python bench.py --steps 1 --warmup 3 --quick --no-overlap"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_${TAG}.csv $B > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
AB_LAUNCHES=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_layer_kernel -s 33 -c 33 -f -o gpurun_out/prof_mlp_${TAG} python tools/mlp_ab.py > gpurun_out/ncu_mlp_${TAG}.log 2>&1
tail -1 gpurun_out/ncu_mlp_${TAG}.log
timeout 300 python tools/mlp_ab.py 2>&1 | grep -v "Warning\|_warn_once" | tail -2
bash tools/sanitizer_r02.sh 2>&1 | grep "^=="
