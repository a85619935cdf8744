#!/usr/bin/env bash
# round-2 first GPU pass: new parity tests, then quick bench variants, then the full bench line
mkdir -p gpurun_out
python -m pytest tests/test_meanshift_gpu.py tests/test_full_size_gpu.py tests/test_reference_dropin_gpu.py tests/test_poses_gpu.py -q -m gpu --timeout 900 2>&1 | tail -60 > gpurun_out/r2a_tests.log
for mode in certified early_exit strict; do
  python bench.py --steps 10 --warmup 3 --quick --ms-mode $mode > gpurun_out/r2a_bench_$mode.json 2> gpurun_out/r2a_bench_$mode.err
done
python bench.py --steps 10 --warmup 3 --quick --no-overlap > gpurun_out/r2a_bench_certified_noov.json 2> gpurun_out/r2a_bench_certified_noov.err
python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench_full.json 2> gpurun_out/r2a_bench_full.err
tail -25 gpurun_out/r2a_tests.log
for f in gpurun_out/r2a_bench_*.json; do echo "== $f"; cut -c1-1500 $f; done
tail -5 gpurun_out/r2a_bench_full.err
