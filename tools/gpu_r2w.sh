#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r02_8gpu.json 2> gpurun_out/bench_r02_8gpu.err
tail -2 gpurun_out/bench_r02_8gpu.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r02_8gpu.json"))
print(round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), d["n_gpus"], d["config"]["parallelism"], d["clocks"])
for k,v in d.get("configs",{}).items(): print(k, round(v["value"],1), round(v["ms_per_step"],3), v["global_batch"], "e2e", round(v["e2e"]["value"],1))
PY
