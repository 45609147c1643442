#!/bin/bash
# 8 GPUs of one box: N = 2 / 4 / 8 gathered-block parity, bench at N = 8 with both exchange modes (config 5 = 10 000 keyframes)
set -u
mkdir -p gpurun_out
export LEVINS_BATCH_CACHE=/tmp/levins_batch_cache
nvidia-smi topo -m > gpurun_out/topo_8.txt 2>&1
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q > gpurun_out/pytest_multi_8.log 2>&1; echo "pytest multi rc=$?"
tail -6 gpurun_out/pytest_multi_8.log
N=8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 --no-stream --no-also --gather peer > gpurun_out/bench_8gpu_peer.json 2> gpurun_out/bench_8gpu_peer.err
echo "bench peer rc=$?"; tail -c 800 gpurun_out/bench_8gpu_peer.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 20 --warmup 3 --no-stream --no-also --no-cpu --no-cfg5 --gather nccl > gpurun_out/bench_8gpu_nccl.json 2> gpurun_out/bench_8gpu_nccl.err
echo "bench nccl rc=$?"; tail -c 400 gpurun_out/bench_8gpu_nccl.err
python - <<'PY'
import json
for m in ('peer','nccl'):
    try:
        d=json.loads(open(f'gpurun_out/bench_8gpu_{m}.json').read().strip().splitlines()[-1])
        print(m, d['n_gpus'], d['value'], d['ms_per_step'], d['e2e'], d['roofline']['launch_ms'])
        print(d.get('cfg5_offline_batch'))
    except Exception as e: print('no json', m, e)
PY
