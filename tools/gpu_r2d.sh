#!/usr/bin/env bash
# look-ahead pipeline + chain slots
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pipeline_gpu.py tests/test_mlp_gpu.py tests/test_pn2_gpu.py tests/test_metrics_gpu.py -q -m gpu -x --timeout 600 2>&1 | tail -15 > gpurun_out/r2d_tests.log
tail -15 gpurun_out/r2d_tests.log
export PVN3D_MLP_CHAIN=0
timeout 300 python bench.py --steps 10 --warmup 3 --quick > gpurun_out/r2d_bench_la.json 2> gpurun_out/r2d_bench_la.err
timeout 300 python bench.py --steps 10 --warmup 3 --quick --no-lookahead > gpurun_out/r2d_bench_nola.json 2> gpurun_out/r2d_bench_nola.err
for s in 2 4 8; do
  PVN3D_MLP_CHAIN=1 PVN3D_CHAIN_SLOTS=$s timeout 300 python bench.py --steps 10 --warmup 3 --quick --no-lookahead > gpurun_out/r2d_bench_chain_s$s.json 2> gpurun_out/r2d_bench_chain_s$s.err
done
PVN3D_MLP_CHAIN=1 PVN3D_CHAIN_SLOTS=4 PVN3D_MLP_TMA=0 timeout 300 python bench.py --steps 10 --warmup 3 --quick --no-lookahead > gpurun_out/r2d_bench_chain_s4_notma.json 2> gpurun_out/r2d_bench_chain_s4_notma.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2d_bench_*.json")):
    try:
        d=json.load(open(f))
        r={x["kernel"][:12]:round(x.get("ms_per_batch",0),3) for x in d.get("rooflines",[])}
        print(f, round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "A", round(d["stage_ms_per_batch"]["hot_path_A_pointnet2msg"],3), r)
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace(".json",".err")).read()[-800:])
PY
