#!/bin/bash
set -u
mkdir -p gpurun_out
export LEVINS_BATCH_CACHE=/tmp/levins_batch_cache
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_c.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_c.log
tail -12 gpurun_out/pytest_c.log
bash scripts/gpu_call_c.sh 2
timeout 600 python le-vins_b200/ab_bench.py product nocull > gpurun_out/ab_c.log 2>&1; cat gpurun_out/ab_c.log
