#!/usr/bin/env bash
# last GPU call of the round: the tree's defaults (8 epilogue warps at one CTA per SM, 64-channel pooled layers on the
# transposed accumulator) against the previous state, whole suite, sanitizer on the layer kernel, full bench line
mkdir -p gpurun_out
ab() { env "$@" AB_LAUNCHES=1 timeout 200 python tools/mlp_ab.py 2>&1 | grep -v "Warning\|_warn_once" | tail -2; }
timeout 900 python -m pytest tests -q -m gpu --timeout 600 2>&1 | tail -4 > gpurun_out/last_tests_on.log; cat gpurun_out/last_tests_on.log
ab PVN3D_MLP_EPI8=1 PVN3D_MLP_POOLT64=1
ab PVN3D_MLP_EPI8=0 PVN3D_MLP_POOLT64=1
ab PVN3D_MLP_EPI8=1 PVN3D_MLP_POOLT64=0
ab PVN3D_MLP_EPI8=0 PVN3D_MLP_POOLT64=0
PVN3D_MLP_EPI8=0 PVN3D_MLP_POOLT64=0 timeout 600 python -m pytest tests/test_mlp_gpu.py tests/test_pipeline_gpu.py tests/test_full_size_gpu.py tests/test_heads_gpu.py -q -m gpu --timeout 300 2>&1 | tail -2 > gpurun_out/last_tests_off.log; cat gpurun_out/last_tests_off.log
CS=/usr/local/cuda/bin/compute-sanitizer
SMALL_MLP='test_dense_layer and (128-32-16 or 300-64-64 or 131-48-80 or 256-32-16 or 1024-64-32 or 160-384-128 or 4144-64-128 or 640-544-256 or 4128-32-64 or 1616-96-64) or test_sa_first or test_fp_first or (test_factored_sa_first_layer_kernels and 1-512-100) or test_factored_fp_first_layer_kernel or (test_factored_fp_layer_channel_major_output and 1-96-40)'
for tool in memcheck racecheck; do
  timeout 600 $CS --tool $tool --print-limit 20 --launch-timeout 0 python -m pytest tests/test_mlp_gpu.py -q -m gpu -x -k "$SMALL_MLP" > gpurun_out/sanitizer_${tool}_mlp_r02.log 2>&1
  echo "== $tool mlp: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' gpurun_out/sanitizer_${tool}_mlp_r02.log | tail -3 | tr '\n' ' ')"
done
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_last.json 2> gpurun_out/bench_last.err
tail -2 gpurun_out/bench_last.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_last.json"))
print(round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "roofline", round(d["roofline"]["frac"],3))
print({k: (round(v["value"],1), round(v["ms_per_step"],2)) for k,v in d["configs"].items()})
print([ (r["kernel"][:20], round(r.get("ms_per_batch",0),3)) for r in d["rooflines"]][:5]); print(d.get("densefusion_heads",{}).get("ms_per_batch"))
PY
