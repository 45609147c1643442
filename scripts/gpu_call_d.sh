#!/bin/bash
set -u
mkdir -p gpurun_out
export LEVINS_BATCH_CACHE=/tmp/levins_batch_cache
timeout 600 python -m pytest tests/test_synthgen.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_d.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_d.log
tail -8 gpurun_out/pytest_d.log
timeout 900 python bench.py --no-also --no-stream > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_d.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_d.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['kernel_time_share'])
print(d['cfg5_offline_batch'])
PY
timeout 900 python le-vins_b200/ab_bench.py product prep4096 prep1024 prepmb1 prep4096mb1 > gpurun_out/ab_d.log 2>&1; cat gpurun_out/ab_d.log
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-stream --no-also --no-cfg5 --e2e-steps 3"
$CMD > gpurun_out/plain_d.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_prepare|k_associate|k_plane" -s 9 -c 3 -o gpurun_out/prof_d $CMD > gpurun_out/ncu_full_d.log 2>&1
tail -3 gpurun_out/ncu_full_d.log
