#!/usr/bin/env bash
mkdir -p gpurun_out
PVN3D_POSE_STREAM=0 timeout 300 python bench.py --steps 10 --warmup 3 --quick > gpurun_out/r2f_bench_la_serialB.json 2> gpurun_out/r2f_bench_la_serialB.err
PVN3D_POSE_STREAM=0 timeout 300 python bench.py --steps 10 --warmup 3 --quick --no-lookahead > gpurun_out/r2f_bench_nola_serialB.json 2> gpurun_out/r2f_bench_nola_serialB.err
PVN3D_POSE_STREAM=0 PVN3D_FPS_CHUNK=32 timeout 300 python bench.py --steps 10 --warmup 3 --quick > gpurun_out/r2f_bench_la_serialB_c32.json 2> gpurun_out/r2f_bench_la_serialB_c32.err
PVN3D_POSE_STREAM=0 PVN3D_FPS_CHUNK=8 timeout 300 python bench.py --steps 10 --warmup 3 --quick > gpurun_out/r2f_bench_la_serialB_c8.json 2> gpurun_out/r2f_bench_la_serialB_c8.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2f_bench_*.json")):
    try:
        d=json.load(open(f))
        print(f, round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "A", round(d["stage_ms_per_batch"]["hot_path_A_pointnet2msg"],3), "B", round(d["stage_ms_per_batch"]["hot_path_B_votes_to_poses"],3))
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace(".json",".err")).read()[-800:])
PY
