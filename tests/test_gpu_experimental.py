"""Kernels that have been written but not yet measured on hardware.  They are OFF by default in the
product; these tests run only with DLRM_EXPERIMENTAL=1 (and a GPU), so the default `pytest -m gpu`
run is unaffected until a round promotes them."""
import os

import numpy as np
import pytest
import torch

from golden_util import Golden

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("DLRM_EXPERIMENTAL") != "1", reason="set DLRM_EXPERIMENTAL=1")]

DEV = "cuda:0"


@pytest.mark.parametrize("name", ["cfg0", "mini_cfg1"])
def test_grouped_wgrad_launch_is_bit_identical(name):
    """dlrm_b200_gemm_tc_run_group: all weight gradients of an MLP in one launch == one launch per layer."""
    from dlrm_b200.engine import Engine, sparse_from_reference

    g = Golden(name)
    X, off, idx, T = g.batch(0)
    sp = sparse_from_reference([torch.from_numpy(o) for o in off], [torch.from_numpy(i) for i in idx], DEV)
    Xd, Td = torch.from_numpy(X).to(DEV), torch.from_numpy(T).to(DEV)
    grads = []
    for grouped in (False, True):
        e = Engine(g.m_spa, g.ln_emb, g.ln_bot, g.ln_top, op=g.op, itself=g.itself, sigmoid_bot=-1,
                   sigmoid_top=len(g.ln_top) - 2, loss=g.loss, loss_threshold=g.thr, device=DEV, max_batch=g.B,
                   gemm="tc")
        e.group_wgrad = grouped
        e.load_params(g.params())
        e.forward(Xd, sp, link=True, skip_head=True)
        e.backward(Xd, sp, Td)
        torch.cuda.synchronize()
        grads.append([e.reduced_dW(w, i).clone() for w in ("bot", "top") for i in range(len(getattr(e, "ln_" + w)) - 1)]
                     + [e.reduced_db(w, i).clone() for w in ("bot", "top") for i in range(len(getattr(e, "ln_" + w)) - 1)])
    for a, b in zip(*grads):
        assert torch.equal(a, b)
