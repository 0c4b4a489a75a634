"""Sharded engine (dlrm_b200.dist.DistEngine: placement, row-split tables, peer-memory exchange, tiny-table path)
against the live-reference goldens: as a 1-rank group on any box, on 2 GPUs when two are visible."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(nproc, port):
    if nproc == 1:
        cmd = [sys.executable, os.path.join(ROOT, "tests", "dist_check.py")]
        env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
        env["MASTER_PORT"] = str(port)
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % nproc,
               "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dist_check.py")]
        env = dict(os.environ)
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    print(r.stdout[-4000:], r.stderr[-3000:])
    assert r.returncode == 0
    assert "PASS" in r.stdout.splitlines()[-1]


def test_sharded_engine_one_rank_matches_reference_goldens():
    """Un-skippable on a 1-GPU box: forced row splits put two shards of a table on the same rank, so the partial-sum
    reduction, multi-destination gradient routes and shard-filtered updates all run."""
    _run(1, 29541)


def test_sharded_engine_two_gpus_matches_reference_goldens():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _run(2, 29543)


def test_cli_under_torchrun_reproduces_the_reference_loss_curve():
    """`torchrun --nproc-per-node 2 dlrm_s_pytorch.py <flags>` (DLRM_Net.distributed_forward behind the reference
    command line): same flags + seed as the single-process reference CLI run recorded in tests/golden/cli_cfg0_A.txt;
    the mean of the ranks' losses is the single-process loss."""
    import re

    import numpy as np

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    flags = open(os.path.join(ROOT, "tests", "golden", "cli_cfg0_A.flags")).read().split()
    want = [float(m.group(1)) for m in re.finditer(r"loss ([0-9.]+)",
                                                   open(os.path.join(ROOT, "tests", "golden", "cli_cfg0_A.txt")).read())]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29547", os.path.join(ROOT, "dlrm_s_pytorch.py"),
           "--arch-sparse-feature-size=16", "--arch-embedding-size=1000-1000-1000", "--arch-mlp-bot=13-512-256-64-16",
           "--arch-mlp-top=512-256-1", "--mini-batch-size=128", "--data-generation=random", "--num-batches=6",
           "--print-freq=1", "--learning-rate=0.1", "--numpy-rand-seed=727", "--use-gpu", "--dist-backend=nccl"] + flags
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT,
                       env=dict(os.environ, DLRM_CLI_GLOBAL_LOSS="1"))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "Using 2 GPU(s)..." in r.stdout
    got = [float(m.group(1)) for m in re.finditer(r"Finished training it \d+/6 of epoch 0, .* loss ([0-9.]+)", r.stdout)]
    assert len(got) == 6, r.stdout
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-5)
