#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_meanshift_gpu.py tests/test_full_size_gpu.py tests/test_poses_gpu.py tests/test_pipeline_gpu.py -q -m gpu --timeout 900 2>&1 | tail -15 > gpurun_out/r2g_tests.log
tail -15 gpurun_out/r2g_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --quick > gpurun_out/r2g_bench_la.json 2> gpurun_out/r2g_bench_la.err
timeout 300 python bench.py --steps 10 --warmup 3 --quick --no-lookahead > gpurun_out/r2g_bench_nola.json 2> gpurun_out/r2g_bench_nola.err
timeout 300 python bench.py --steps 10 --warmup 3 --quick --config ycb > gpurun_out/r2g_bench_ycb_la.json 2> gpurun_out/r2g_bench_ycb_la.err
timeout 300 python bench.py --steps 10 --warmup 3 --quick --config ycb --no-lookahead > gpurun_out/r2g_bench_ycb_nola.json 2> gpurun_out/r2g_bench_ycb_nola.err
PVN3D_FPS_CHUNK=8 timeout 300 python bench.py --steps 10 --warmup 3 --quick --config ycb > gpurun_out/r2g_bench_ycb_la_c8.json 2> gpurun_out/r2g_bench_ycb_la_c8.err
PVN3D_FPS_CHUNK=12 timeout 300 python bench.py --steps 10 --warmup 3 --quick > gpurun_out/r2g_bench_la_c12.json 2> gpurun_out/r2g_bench_la_c12.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2g_bench_*.json")):
    try:
        d=json.load(open(f))
        print(f, round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "A", round(d["stage_ms_per_batch"]["hot_path_A_pointnet2msg"],3), "B", round(d["stage_ms_per_batch"]["hot_path_B_votes_to_poses"],3), d.get("meanshift_certified_fits"))
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace(".json",".err")).read()[-800:])
PY
