#!/bin/bash
# final single-GPU validation of the tree: all GPU tests, the default bench line, the launch list of the same command
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_final.log
tail -6 gpurun_out/pytest_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_final.log
timeout 600 python bench.py --impl reference > gpurun_out/bench_final_reference.json 2> gpurun_out/bench_final_reference.err; echo "reference rc=$?"
timeout 1200 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_final.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1])
r=json.loads(open('gpurun_out/bench_final_reference.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'ref', r['value'], 'same config', d['config']==r['config'])
print(d['roofline']); print(d['parity']); print(d['streaming']); print(d['also']); print(d['cfg5_offline_batch'])
PY
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-stream --no-also --no-cfg5 --e2e-steps 3"
$CMD > gpurun_out/plain_final.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_final.csv $CMD > gpurun_out/ncu_launch_final.log 2>&1
