#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mlp_gpu.py tests/test_full_size_gpu.py -q -m gpu --timeout 600 -k "not meanshift" 2>&1 | tail -8 > gpurun_out/r2h_tests.log
tail -8 gpurun_out/r2h_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --quick > gpurun_out/r2h_bench_occ2.json 2> gpurun_out/r2h_bench_occ2.err
PVN3D_MLP_OCC=1 timeout 300 python bench.py --steps 10 --warmup 3 --quick > gpurun_out/r2h_bench_occ1.json 2> gpurun_out/r2h_bench_occ1.err
timeout 300 python bench.py --steps 10 --warmup 3 --quick --config ycb > gpurun_out/r2h_bench_ycb_occ2.json 2> gpurun_out/r2h_bench_ycb_occ2.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r02h_occ2.csv python bench.py --steps 1 --warmup 3 --quick --no-overlap > gpurun_out/bench_under_ncu_r02h.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2h_bench_*.json")):
    try:
        d=json.load(open(f))
        r={x["kernel"][:12]:round(x.get("ms_per_batch",0),3) for x in d.get("rooflines",[])}
        print(f, round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "A", round(d["stage_ms_per_batch"]["hot_path_A_pointnet2msg"],3), r)
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace(".json",".err")).read()[-800:])
PY
