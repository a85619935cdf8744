#!/usr/bin/env bash
# round-2 first GPU pass.  Everything runs under `timeout`: a pipeline bug in a new kernel hangs, it does not fail.
mkdir -p gpurun_out
# 1. chain kernel bring-up
timeout 900 python -m pytest tests/test_mlp_gpu.py -q -m gpu -x --timeout 300 -k "chain" 2>&1 | tail -30 > gpurun_out/r2a_chain_tests.log
CHAIN_OK=0; grep -q " passed" gpurun_out/r2a_chain_tests.log && ! grep -q "failed\|error\|Timeout" gpurun_out/r2a_chain_tests.log && CHAIN_OK=1
echo "chain ok: $CHAIN_OK" >> gpurun_out/r2a_chain_tests.log
# 2. parity at full size + drop-in, per-layer engine (known good)
export PVN3D_MLP_CHAIN=0
timeout 1500 python -m pytest tests/test_meanshift_gpu.py tests/test_full_size_gpu.py tests/test_reference_dropin_gpu.py tests/test_poses_gpu.py -q -m gpu --timeout 900 2>&1 | tail -60 > gpurun_out/r2a_tests.log
for mode in certified early_exit strict; do
  timeout 300 python bench.py --steps 10 --warmup 3 --quick --ms-mode $mode > gpurun_out/r2a_bench_$mode.json 2> gpurun_out/r2a_bench_$mode.err
done
timeout 300 python bench.py --steps 10 --warmup 3 --quick --no-overlap > gpurun_out/r2a_bench_certified_noov.json 2> gpurun_out/r2a_bench_certified_noov.err
if [ $CHAIN_OK = 1 ]; then
  PVN3D_MLP_CHAIN=1 timeout 300 python bench.py --steps 10 --warmup 3 --quick > gpurun_out/r2a_bench_chain.json 2> gpurun_out/r2a_bench_chain.err
fi
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench_full.json 2> gpurun_out/r2a_bench_full.err
tail -12 gpurun_out/r2a_chain_tests.log
tail -25 gpurun_out/r2a_tests.log
for f in gpurun_out/r2a_bench_*.json; do echo "== $f"; cut -c1-1800 $f; done
tail -5 gpurun_out/r2a_bench_full.err
