#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pipeline_gpu.py tests/test_mlp_gpu.py tests/test_pn2_gpu.py tests/test_metrics_gpu.py -q -m gpu --timeout 600 2>&1 | tail -15 > gpurun_out/r2e_tests.log
tail -15 gpurun_out/r2e_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --quick > gpurun_out/r2e_bench_la.json 2> gpurun_out/r2e_bench_la.err
timeout 300 python bench.py --steps 10 --warmup 3 --quick --no-lookahead > gpurun_out/r2e_bench_nola.json 2> gpurun_out/r2e_bench_nola.err
PVN3D_MLP_TMA=0 timeout 300 python bench.py --steps 10 --warmup 3 --quick --no-lookahead > gpurun_out/r2e_bench_nola_notma.json 2> gpurun_out/r2e_bench_nola_notma.err
timeout 300 python bench.py --steps 10 --warmup 3 --quick --config ycb > gpurun_out/r2e_bench_ycb_la.json 2> gpurun_out/r2e_bench_ycb_la.err
timeout 300 python bench.py --steps 10 --warmup 3 --quick --config ycb --no-lookahead > gpurun_out/r2e_bench_ycb_nola.json 2> gpurun_out/r2e_bench_ycb_nola.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2e_bench_*.json")):
    try:
        d=json.load(open(f))
        r={x["kernel"][:12]:round(x.get("ms_per_batch",0),3) for x in d.get("rooflines",[])}
        print(f, round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "A", round(d["stage_ms_per_batch"]["hot_path_A_pointnet2msg"],3), "B", round(d["stage_ms_per_batch"]["hot_path_B_votes_to_poses"],3), r)
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace(".json",".err")).read()[-800:])
PY
