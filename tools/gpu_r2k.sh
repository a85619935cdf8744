#!/usr/bin/env bash
# 2-GPU run of the bench (driver's launch line) + gloo/nccl sanity
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2k_bench_2gpu.json 2> gpurun_out/r2k_bench_2gpu.err
tail -3 gpurun_out/r2k_bench_2gpu.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2k_bench_2gpu.json"))
print(round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), d["n_gpus"], d["config"]["parallelism"])
for k,v in d.get("configs",{}).items(): print(k, round(v["value"],1), round(v["ms_per_step"],3), v["global_batch"])
PY
