#!/usr/bin/env bash
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 20 --warmup 5 --quick "$@" > gpurun_out/r2l_$tag.json 2> gpurun_out/r2l_$tag.err; }
run la
run nola --no-lookahead
run noov --no-overlap
run la2
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2l_*.json")):
    try:
        d=json.load(open(f)); print(f, round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), round(d["e2e"]["ms_per_step"],3))
    except Exception as e: print(f, "ERR", e)
PY
