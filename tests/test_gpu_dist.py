"""2-GPU parity of the sharded engine (skipped unless >= 2 devices are visible)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("exchange", ["nccl", "p2p"])
@pytest.mark.parametrize("gemm", ["tc"])
def test_two_gpu_sharded_matches_single_device(gemm, exchange):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, DLRM_GEMM=gemm, DLRM_EXCHANGE=exchange)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29517",
                        os.path.join(ROOT, "tests", "dist_check.py")], env=env, capture_output=True, text=True,
                       timeout=600)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0
