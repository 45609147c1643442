#!/bin/bash
# N GPUs of one box: gathered block == single-GPU results (peer mode and NCCL), then the bench at N with both exchange modes
set -u
N=${1:-2}
mkdir -p gpurun_out
export LEVINS_BATCH_CACHE=/tmp/levins_batch_cache
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -x > gpurun_out/pytest_multi_$N.log 2>&1; echo "pytest multi rc=$?"
tail -15 gpurun_out/pytest_multi_$N.log
for mode in peer nccl; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu --no-stream --no-also --gather $mode > gpurun_out/bench_${N}gpu_$mode.json 2> gpurun_out/bench_${N}gpu_$mode.err
  echo "bench $mode rc=$?"; tail -c 600 gpurun_out/bench_${N}gpu_$mode.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_${N}gpu_$mode.json').read().strip().splitlines()[-1])
    print('$mode', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e'], d['roofline']['launch_ms'])
except Exception as e: print('no json', e)
PY
done
