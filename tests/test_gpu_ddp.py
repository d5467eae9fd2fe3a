"""Multi-GPU numerics (needs >= 2 GPUs: ``gpurun --gpus 2 -- python -m pytest tests/test_gpu_ddp.py -m gpu``)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("config", ["ZINC", "PATTERN"])
def test_two_ranks_equal_one_process_on_the_union_batch(cuda, config):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, DDP2_CONFIG=config)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29631",
                        os.path.join(ROOT, "tests", "ddp_two_rank_check.py")],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "DDP2 OK" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])
