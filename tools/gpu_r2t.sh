#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_pn2_gpu.py tests/test_mlp_gpu.py tests/test_full_size_gpu.py tests/test_pipeline_gpu.py tests/test_reference_dropin_gpu.py -q -m gpu --timeout 900 2>&1 | tail -12 > gpurun_out/r2t_tests.log
tail -12 gpurun_out/r2t_tests.log
timeout 300 python bench.py --steps 20 --warmup 5 --quick > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err
PVN3D_NN_SLAB=0 timeout 300 python bench.py --steps 20 --warmup 5 --quick > gpurun_out/r2t_bench_noslab.json 2> gpurun_out/r2t_bench_noslab.err
timeout 300 python bench.py --steps 20 --warmup 5 --quick --config ycb > gpurun_out/r2t_bench_ycb.json 2> gpurun_out/r2t_bench_ycb.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2t_bench*.json")):
    try:
        d=json.load(open(f))
        r={x["kernel"][:12]:round(x.get("ms_per_batch",0),3) for x in d.get("rooflines",[])}
        print(f, round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "A", round(d["stage_ms_per_batch"]["hot_path_A_pointnet2msg"],3), "B", round(d["stage_ms_per_batch"]["hot_path_B_votes_to_poses"],3), r)
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
