#!/usr/bin/env bash
# Run ON THE GPU BOX (under gpurun): bench line + ncu launch list + full captures of the two top kernels.
# Outputs land in gpurun_out/ (scratch); summaries are copied into profiles/ by tools/summarise_ncu.py.
set -u
mkdir -p gpurun_out
TAG=${1:-r01}
BENCH_ARGS=${BENCH_ARGS:---steps 2 --warmup 3}
echo "== bench ($BENCH_ARGS)"
timeout 900 python bench.py $BENCH_ARGS > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -c 600 gpurun_out/bench_${TAG}.err
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline \
    > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
echo "== ncu full: mlp_layer_kernel (32 launches per step: skip 3 warm-up steps)"
timeout 900 ncu --set full --clock-control none -k regex:mlp_layer_kernel -s 96 -c 32 \
    -f -o gpurun_out/prof_mlp_${TAG} python bench.py --steps 1 --warmup 3 --no-cpu-baseline \
    > gpurun_out/ncu_mlp_${TAG}.log 2>&1
echo "== ncu full: ball_scan_kernel + group_write_kernel (tools/qg_roofline.py)"
timeout 900 ncu --set full --clock-control none -k "regex:ball_scan_kernel|group_write_kernel" -c 8 \
    -f -o gpurun_out/prof_qg_${TAG} python tools/qg_roofline.py \
    > gpurun_out/ncu_qg_${TAG}.log 2>&1
echo "== ncu full: ms_iterate_kernel"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:ms_iterate_kernel -s 6 -c 2 \
    -f -o gpurun_out/prof_ms_${TAG} python bench.py --steps 1 --warmup 3 --no-cpu-baseline \
    > gpurun_out/ncu_ms_${TAG}.log 2>&1
gzip -f gpurun_out/*.ncu-rep   # merged back only if gpurun_out/ stays under 64 MiB
ls -la gpurun_out
