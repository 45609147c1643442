"""N > 1 on hardware: the gathered block (peer mode and NCCL, both through the C ABI) equals the single-GPU results bytewise.
Skipped on boxes with one GPU (the driver's round-end GPU tier); run with `gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    try:
        import torch

        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4, 8])
def test_gathered_block_equals_single_gpu(built_lib, world):
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
                        os.path.join(ROOT, "tests", "multi_gpu_driver.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and f"MULTI_GPU_OK {world}" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
