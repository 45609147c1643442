#!/bin/bash
# validate the round-2 additions on one GPU: all GPU tests (verbose failures), a default bench run
set -u
mkdir -p gpurun_out
export LEVINS_BATCH_CACHE=/tmp/levins_batch_cache
timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/pytest_b.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_b.log
tail -25 gpurun_out/pytest_b.log
timeout 900 python bench.py > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; echo "bench rc=$?"
tail -c 400 gpurun_out/bench_b.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_b.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','streaming','e2e','parity')})
print(d['roofline'])
PY
