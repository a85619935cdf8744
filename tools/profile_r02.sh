#!/usr/bin/env bash
# Run ON THE GPU BOX (under gpurun): round-2 ncu evidence.  Launch list of one serialised step + full captures of the
# kernels on the step.  Outputs land in gpurun_out/ (scratch); tools/summarise_ncu.py r02 copies summaries to profiles/.
set -u
mkdir -p gpurun_out
TAG=${1:-r02}
B="python bench.py --steps 1 --warmup 3 --quick --no-overlap"
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_${TAG}.csv $B > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
echo "== ncu full: mlp_layer_kernel (32 launches of one forward; skip the forwards of the warm-up)"
# one forward = 33 mlp_layer_kernel launches (8 SA scales x (U, L2', L3) + FP4-2 x 2 + FP1 x 3); bench --quick runs 5 warm-up
# steps + 1 timed + profile passes before the stage split: skip 6 forwards
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlp_layer_kernel -s 198 -c 33 -f -o gpurun_out/prof_mlp_${TAG} $B > gpurun_out/ncu_mlp_${TAG}.log 2>&1
echo "== ncu full: ms_density_kernel, ms_witness_kernel, ms_fallback_kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:ms_density_pruned_kernel" -s 8 -c 2 -f -o gpurun_out/prof_msdens_${TAG} $B > gpurun_out/ncu_msdens_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:ms_witness_kernel|ms_fallback_kernel" -s 8 -c 2 -f -o gpurun_out/prof_mswit_${TAG} $B > gpurun_out/ncu_mswit_${TAG}.log 2>&1
echo "== ncu full: fps_regs_kernel, three_nn_kernel, ball_scan_kernel (one forward)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fps_regs_kernel -s 16 -c 4 -f -o gpurun_out/prof_fps_${TAG} $B > gpurun_out/ncu_fps_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:three_nn_kernel|three_nn_slab_kernel|nn_sort_known_kernel" -s 28 -c 7 -f -o gpurun_out/prof_nn_${TAG} $B > gpurun_out/ncu_nn_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ball_scan_kernel -s 16 -c 4 -f -o gpurun_out/prof_ball_${TAG} $B > gpurun_out/ncu_ball_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none -k "regex:sa_factor_table_kernel|sa_centre_term_kernel|transpose_kernel|gather_xyz_kernel" -s 60 -c 10 -f -o gpurun_out/prof_glue_${TAG} $B > gpurun_out/ncu_glue_${TAG}.log 2>&1
echo "== strict mode: ms_iterate_kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ms_iterate_kernel -s 4 -c 1 -f -o gpurun_out/prof_ms_${TAG} $B --ms-mode strict > gpurun_out/ncu_ms_${TAG}.log 2>&1
gzip -f gpurun_out/*_${TAG}.ncu-rep
ls -la gpurun_out | grep ${TAG}
