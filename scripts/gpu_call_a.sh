#!/bin/bash
# first GPU call of round 2: tests, bench on cfg3, cull / wave A/B, launch list + one full capture of k_deskew_bin
set -u
mkdir -p gpurun_out
export LEVINS_BATCH_CACHE=/tmp/levins_batch_cache
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/clocks_a.csv 2>/dev/null &
SMI=$!
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_a.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_a.log
tail -5 gpurun_out/pytest_a.log
timeout 900 python bench.py > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench_a.json
timeout 900 python le-vins_b200/ab_bench.py product nocull product:--wave=8 product:--wave=16 > gpurun_out/ab_a.log 2>&1
cat gpurun_out/ab_a.log
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-stream --no-also --e2e-steps 3"
$CMD > gpurun_out/plain_a.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_a.csv $CMD > gpurun_out/ncu_launch_a.log 2>&1
$CMD > gpurun_out/plain_a2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_deskew_bin -s 3 -c 2 -o gpurun_out/prof_a $CMD > gpurun_out/ncu_full_a.log 2>&1
kill $SMI 2>/dev/null
ls -la gpurun_out | tail -20
