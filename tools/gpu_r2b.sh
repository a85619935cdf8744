#!/usr/bin/env bash
# chain kernel bring-up: its tests under a hard timeout (a pipeline bug hangs, it does not fail)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_mlp_gpu.py -q -m gpu -x --timeout 300 -k "chain" 2>&1 | tail -30 > gpurun_out/r2b_chain_tests.log
tail -30 gpurun_out/r2b_chain_tests.log
