#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 2>&1 | tail -12 > gpurun_out/r2i_tests.log
tail -12 gpurun_out/r2i_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --quick > gpurun_out/r2i_bench_a.json 2> gpurun_out/r2i_bench_a.err
PVN3D_MLP_ROUND_TABLES=0 timeout 300 python bench.py --steps 10 --warmup 3 --quick > gpurun_out/r2i_bench_noround.json 2> gpurun_out/r2i_bench_noround.err
timeout 300 python bench.py --steps 10 --warmup 3 --quick --no-lookahead > gpurun_out/r2i_bench_nola.json 2> gpurun_out/r2i_bench_nola.err
timeout 300 python bench.py --steps 10 --warmup 3 --quick --config ycb > gpurun_out/r2i_bench_ycb.json 2> gpurun_out/r2i_bench_ycb.err
timeout 300 python bench.py --steps 20 --warmup 5 --quick > gpurun_out/r2i_bench_b.json 2> gpurun_out/r2i_bench_b.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2i_bench_*.json")):
    try:
        d=json.load(open(f))
        r={x["kernel"][:12]:round(x.get("ms_per_batch",0),3) for x in d.get("rooflines",[])}
        print(f, round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "A", round(d["stage_ms_per_batch"]["hot_path_A_pointnet2msg"],3), "B", round(d["stage_ms_per_batch"]["hot_path_B_votes_to_poses"],3), r)
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace(".json",".err")).read()[-800:])
PY
