#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_meanshift_gpu.py tests/test_full_size_gpu.py tests/test_poses_gpu.py tests/test_pipeline_gpu.py tests/test_reference_dropin_gpu.py -q -m gpu --timeout 900 2>&1 | tail -15 > gpurun_out/r2s_tests.log
tail -15 gpurun_out/r2s_tests.log
timeout 300 python bench.py --steps 20 --warmup 5 --quick > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err
PVN3D_MS_PRUNED_DENSITY=0 timeout 300 python bench.py --steps 20 --warmup 5 --quick > gpurun_out/r2s_bench_brute.json 2> gpurun_out/r2s_bench_brute.err
timeout 300 python bench.py --steps 20 --warmup 5 --quick --config ycb > gpurun_out/r2s_bench_ycb.json 2> gpurun_out/r2s_bench_ycb.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2s_bench*.json")):
    try:
        d=json.load(open(f))
        print(f, round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "A", round(d["stage_ms_per_batch"]["hot_path_A_pointnet2msg"],3), "B", round(d["stage_ms_per_batch"]["hot_path_B_votes_to_poses"],3), d.get("meanshift_certified_fits",{}).get("certified"))
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
