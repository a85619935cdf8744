#!/usr/bin/env bash
# ncu launch lists: per-layer engine vs chained engine (quick bench, 1 timed step)
mkdir -p gpurun_out
for c in 0 1; do
  PVN3D_MLP_CHAIN=$c timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
    --log-file gpurun_out/launches_r02c_chain$c.csv python bench.py --steps 1 --warmup 3 --quick --no-overlap \
    > gpurun_out/bench_under_ncu_r02c_chain$c.log 2>&1
done
# full capture of the 12 chained launches of one forward (skip the warm-up forwards: 4 steps x 12 = 48 ... profile() adds more; take the first 12 after 60)
PVN3D_MLP_CHAIN=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlp_chain_kernel -s 60 -c 12 \
    -f -o gpurun_out/prof_chain_r02c python bench.py --steps 1 --warmup 3 --quick --no-overlap > gpurun_out/ncu_chain_r02c.log 2>&1
gzip -f gpurun_out/prof_chain_r02c.ncu-rep
ls -la gpurun_out | tail -8
