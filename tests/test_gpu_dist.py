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
