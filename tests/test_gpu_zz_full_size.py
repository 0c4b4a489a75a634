"""Full-size checks that need most of a GPU's memory and a few seconds each; collected last (file name) so that a
problem here cannot hide the rest of the suite behind `pytest -x`."""
import numpy as np
import pytest
import torch

from golden_util import O

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def test_cfg2_full_size_one_training_step_matches_oracle():
    """BASELINE.json configs[2] at FULL size (26 x 1e6 x 128 tables, bot 13-512-256-128, top 479-1024-512-256-1,
    B = 2048), one fwd + bwd + RWSAdagrad step on the product path (tensor-core GEMMs) against the oracle.  The
    oracle runs on COMPACTED tables: only the rows the batch touches exist there (gathered from the device before
    the step, indices renumbered) -- untouched rows do not enter the step, and that they did not change on the device
    is checked with per-table float64 checksums."""
    from dlrm_b200.data import make_batch, to_device_packed
    from dlrm_b200.engine import Engine

    T, R, D, B, lr = 26, 1_000_000, 128, 2048, 0.01
    ln_bot, ln_top = [13, 512, 256, D], [D + (T + 1) * T // 2, 1024, 512, 256, 1]
    e = Engine(D, [R] * T, ln_bot, ln_top, loss="bce", sigmoid_top=len(ln_top) - 2, device=DEV, max_batch=B, gemm="tc")
    e.init_params(7)
    e.ensure_optimizer_state("rwsadagrad")
    hb = make_batch(np.random.default_rng(11), [R] * T, B, 13, lmax=10)
    db = to_device_packed(hb, DEV)
    rng = np.random.default_rng(12)
    X = rng.random((B, 13), dtype=np.float32)
    tgt = np.round(rng.random((B, 1), dtype=np.float32))
    # compacted oracle model
    uniq, lS_o, lS_i, emb = [], [], [], []
    for k in range(T):
        idx = hb.indices[hb.offsets[k, 0]:hb.offsets[k, B]]
        u, inv = np.unique(idx, return_inverse=True)
        uniq.append(u)
        lS_i.append(inv.astype(np.int64))
        lS_o.append((hb.offsets[k, :B] - hb.offsets[k, 0]).astype(np.int64))
        emb.append(e.table(k)[torch.from_numpy(u).to(DEV)].cpu().numpy())
    params = dict(emb=emb, v_W_l=None,
                  bot=[(e.W["bot"][i].cpu().numpy().copy(), e.b["bot"][i].cpu().numpy().copy()) for i in range(3)],
                  top=[(e.W["top"][i].cpu().numpy().copy(), e.b["top"][i].cpu().numpy().copy()) for i in range(4)])
    dev_u = [torch.from_numpy(u).to(DEV) for u in uniq]
    # float64 checksum of the UNTOUCHED rows of every table (whole table minus the touched rows, both on the device)
    rest0 = [float(e.table(k).double().sum().item()) - float(e.table(k)[dev_u[k]].double().sum().item()) for k in range(T)]
    state = O.new_state(params)
    r = O.train_step(params, state, X, lS_o, lS_i, tgt, lr=lr, optimizer="rwsadagrad", loss="bce",
                     sigmoid_top=len(ln_top) - 2)
    loss = float(e.train_step(torch.from_numpy(X).to(DEV), db.sparse, torch.from_numpy(tgt).to(DEV), lr,
                              optimizer="rwsadagrad").item())
    torch.cuda.synchronize()
    assert abs(loss - float(r["loss"])) < 1e-5
    errs, merr = [], []
    for k in range(T):
        u = torch.from_numpy(uniq[k]).to(DEV)
        got = e.table(k)[u].cpu().numpy()
        errs.append(np.abs(got - params["emb"][k]).ravel())
        m = e.momentum[int(e.row_base[k]):int(e.row_base[k + 1])]
        gm = m[u].cpu().numpy()
        # rows whose gradient is at the GEMMs' rounding level have O(1) relative error in mean(g^2): scale those by
        # the table's typical accumulator instead
        merr.append(np.abs(gm - state["mom"][k]) / np.maximum(state["mom"][k], 1e-2 * np.median(state["mom"][k])))
        # untouched rows: accumulators still zero, checksum unchanged
        assert int((m != 0).sum().item()) <= uniq[k].size
        rest1 = float(e.table(k).double().sum().item()) - float(e.table(k)[u].double().sum().item())
        assert abs(rest1 - rest0[k]) < 1e-4, (k, rest1, rest0[k])      # one changed row would move it by ~0.1
    errs, merr = np.concatenate(errs), np.concatenate(merr)
    # the first Adagrad step moves every element by lr * g / |g|_rms: errors are relative errors of g times lr
    print("cfg2 full size: row err median %.3g p999 %.3g max %.3g; accumulator rel err median %.3g p999 %.3g"
          % (np.median(errs), np.quantile(errs, 0.999), errs.max(), np.median(merr), np.quantile(merr, 0.999)))
    assert np.median(errs) < 1e-6 and np.quantile(errs, 0.999) < 2.5 * lr * 1e-2
    assert np.median(merr) < 1e-3 and np.quantile(merr, 0.999) < 0.2
    for nm in ("bot", "top"):
        for i, (W, b) in enumerate(params[nm]):
            dw = np.abs(e.W[nm][i].cpu().numpy() - W)
            assert np.median(dw) < 2e-5 and np.quantile(dw, 0.999) < 2.5 * lr, (nm, i, np.median(dw), dw.max())
    assert int(e.head.abs().sum().item()) == 0
