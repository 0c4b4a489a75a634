"""dlrm_b200.cli.LRPolicy against the reference's LRPolicyScheduler (dlrm_s_pytorch.py:169-203):
live when /root/reference is present, and against values recorded from it otherwise.  CPU only."""
import os
import sys

import pytest
import torch

from dlrm_b200.cli import LRPolicy

REF = os.environ.get("DLRM_REFERENCE", "/root/reference")

# (warmup, decay_start, decay_steps) -> learning rate seen by steps 1..14 with base lr 0.5,
# recorded from the live reference (optimizer.step(); scheduler.step() per iteration)
RECORDED = {
    (4, 6, 5): [0.125, 0.25, 0.375, 0.375, 0.375, 0.5, 0.32000000000000006, 0.18, 0.08000000000000002,
                0.020000000000000004, 0.020000000000000004, 0.020000000000000004, 0.020000000000000004,
                0.020000000000000004],
    (0, 0, 0): [0.5] * 14,
    (3, 3, 4): [0.16666666666666669, 0.33333333333333337, 0.5, 0.28125, 0.125, 0.03125, 0.03125, 0.03125, 0.03125,
                0.03125, 0.03125, 0.03125, 0.03125, 0.03125],
    (10, 12, 30): [0.04999999999999999, 0.09999999999999998, 0.15000000000000002, 0.2, 0.25, 0.3, 0.35, 0.4, 0.45,
                   0.45, 0.45, 0.5, 0.4672222222222222, 0.4355555555555556],
}


def _run(make_sched, n=14, lr=0.5):
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([p], lr=lr)
    sched = make_sched(opt)
    seen = []
    for _ in range(n):
        seen.append(opt.param_groups[0]["lr"])
        opt.step()
        sched.step()
    return seen


@pytest.mark.parametrize("cfg", sorted(RECORDED))
def test_lr_policy_recorded(cfg):
    got = _run(lambda o: LRPolicy(o, *cfg))
    assert got == RECORDED[cfg]


@pytest.mark.skipif(not os.path.isdir(REF), reason="live reference not present (GPU box)")
@pytest.mark.parametrize("cfg", [(4, 6, 5), (0, 0, 0), (3, 3, 4), (10, 12, 30), (1, 1, 1)])
def test_lr_policy_live_reference(cfg, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    import builtins
    import importlib.util

    monkeypatch.syspath_prepend(REF)        # the reference's own imports (dlrm_data_pytorch, ...) resolve there
    keep = builtins.print
    try:                                    # loaded under another name: this repo ships a dlrm_s_pytorch.py too
        spec = importlib.util.spec_from_file_location("ref_dlrm_s_pytorch", os.path.join(REF, "dlrm_s_pytorch.py"))
        R = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(R)
    finally:
        builtins.print = keep
    want = _run(lambda o: R.LRPolicyScheduler(o, *cfg), n=50)
    got = _run(lambda o: LRPolicy(o, *cfg), n=50)
    assert got == want
