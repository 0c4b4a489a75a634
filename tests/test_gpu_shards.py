"""Sharded-placement machinery on ONE GPU: row-split tables (partial sums + reduction), tiny-table dense update,
device-side multi-hot generator, index range check, block copy -- every kernel of the multi-GPU path that does not
need a second device, against the numpy oracle (oracle/dlrm_numpy.py)."""
import copy
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _sp(off, idx):
    from dlrm_b200.engine import sparse_from_reference

    return sparse_from_reference([torch.from_numpy(o) for o in off], [torch.from_numpy(i) for i in idx], DEV)


def _model(rng, D, ln_emb, ln_bot, tail, B, lmax):
    from oracle import dlrm_numpy as O

    F = len(ln_emb) + 1
    ln_top = [D + F * (F - 1) // 2] + tail
    params = O.random_params(rng, D, ln_emb, ln_bot, ln_top)
    X, off, idx = O.random_batch(rng, ln_emb, B, ln_bot[0], lmax)
    tgt = np.round(rng.random((B, 1))).astype(np.float32)
    return ln_top, params, X, off, idx, tgt


@pytest.mark.parametrize("gemm", ["tc", "simt"])
def test_row_split_tables_on_one_gpu_match_oracle(gemm):
    """Tables 1 and 3 stored as two row-range shards each (placement.plan(force_split=...) at world 1): the gather
    pools per shard, reduce_partials adds them, the update touches every shard's rows.  Forward within 1e-5 of
    the oracle (fp32 add order differs from one sequential sum), training steps like the unsplit engine."""
    from oracle import dlrm_numpy as O
    from dlrm_b200 import placement as P, sharding as S
    from dlrm_b200.engine import Engine

    rng = np.random.default_rng(5)
    D, ln_emb, ln_bot, tail, B = 64, [3000, 777, 40, 1501], [13, 64, 64], [64, 32, 1], 200
    ln_top, params, X, off, idx, tgt = _model(rng, D, ln_emb, ln_bot, tail, B, 9)
    pl = P.plan(ln_emb, [5.0] * 4, 1, force_split=[1, 3])
    assert pl.split_tables() == [1, 3] and len(pl.of_rank(0)) == 6
    kw = S.engine_kwargs(pl, 0, len(ln_emb))
    e = Engine(D, kw["ln_emb"], ln_bot, ln_top, loss="bce", sigmoid_top=len(ln_top) - 2, device=DEV, max_batch=B,
               gemm=gemm, shards=kw["shards"], split_slots=kw["split_slots"], n_features=kw["n_features"])
    e.load_params(S.slice_params(params, pl, 0))
    streams = S.local_streams(list(zip(off, idx)), pl, 0)
    sp = _sp([o for o, _ in streams], [i for _, i in streams])
    Xd, Td = torch.from_numpy(X).to(DEV), torch.from_numpy(tgt).to(DEV)
    want = O.dlrm_forward(params, X, off, idx)
    got = e.forward(Xd, sp).cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-5)
    pc, st = copy.deepcopy(params), None
    st = O.new_state(pc)
    for step in range(2):
        loss = float(e.train_step(Xd, sp, Td, 0.05, "rwsadagrad").item())
        r = O.train_step(pc, st, X, off, idx, tgt, lr=0.05, optimizer="rwsadagrad", loss="bce")
        assert abs(loss - float(r["loss"])) < 2e-5, (step, loss, float(r["loss"]))
    e.lib.dlrm_b200_check_device_errors(None)
    for j, s in enumerate(pl.of_rank(0)):
        Wj = e.table(j).cpu().numpy()
        ref = pc["emb"][s.table][s.row_lo:s.row_hi]
        err = np.abs(Wj - ref)
        assert np.median(err) < 1e-6 and np.quantile(err, 0.999) < 5e-3, (j, err.max())
        mj = e.momentum[int(e.row_base[j]):int(e.row_base[j + 1])].cpu().numpy()
        np.testing.assert_allclose(mj, st["mom"][s.table][s.row_lo:s.row_hi], rtol=5e-3, atol=1e-9)
    assert int(e.head.abs().sum().item()) == 0


def test_remote_read_forward_is_bit_exact_and_trains_like_the_partial_sum_variant():
    """Row-split tables, remote-read forward (here every 'remote' shard is a local one): the sample owner pools the
    whole bag in index order from the shards -> the pooled rows equal the reference's EmbeddingBag bit for bit;
    the training steps (stand-alone link of the shard's occurrences + the usual update) match the oracle."""
    from oracle import dlrm_numpy as O
    from dlrm_b200 import placement as P, sharding as S
    from dlrm_b200.engine import Engine

    rng = np.random.default_rng(6)
    D, ln_emb, ln_bot, tail, B = 128, [3000, 777, 40, 1501], [13, 64, 128], [64, 32, 1], 200
    ln_top, params, X, off, idx, tgt = _model(rng, D, ln_emb, ln_bot, tail, B, 9)
    pl = P.plan(ln_emb, [5.0] * 4, 1, force_split=[1, 3])
    kw = S.engine_kwargs(pl, 0, len(ln_emb))
    e = Engine(D, kw["ln_emb"], ln_bot, ln_top, loss="bce", sigmoid_top=len(ln_top) - 2, device=DEV, max_batch=B,
               gemm="tc", shards=kw["shards"], split_slots=kw["split_slots"], n_features=kw["n_features"])
    e.use_remote_reads()
    e.load_params(S.slice_params(params, pl, 0))
    streams = S.local_streams(list(zip(off, idx)), pl, 0)
    sp = _sp([o for o, _ in streams], [i for _, i in streams])
    Xd, Td = torch.from_numpy(X).to(DEV), torch.from_numpy(tgt).to(DEV)
    got = e.forward(Xd, sp).cpu().numpy()
    np.testing.assert_allclose(got, O.dlrm_forward(params, X, off, idx), rtol=0, atol=1e-5)
    Tb = e.Tbuf[:B].cpu().numpy()
    for t in range(len(ln_emb)):
        assert np.array_equal(Tb[:, 1 + t, :], O.emb_bag_sum(params["emb"][t], idx[t], off[t])), t
    pc = copy.deepcopy(params)
    st = O.new_state(pc)
    for step in range(2):
        loss = float(e.train_step(Xd, sp, Td, 0.05, "rwsadagrad").item())
        r = O.train_step(pc, st, X, off, idx, tgt, lr=0.05, optimizer="rwsadagrad", loss="bce")
        assert abs(loss - float(r["loss"])) < 2e-5, (step, loss, float(r["loss"]))
    assert e.lib.dlrm_b200_check_device_errors(None) == 0
    assert int(e.head.abs().sum().item()) == 0


@pytest.mark.parametrize("opt", ["rwsadagrad", "sgd"])
@pytest.mark.parametrize("idx_dtype", [np.int64, np.int32])
def test_tiny_tables_dense_update_matches_oracle(opt, idx_dtype):
    """3 / 10 / 155-row tables at batch 4096 (every row hit hundreds of times): the dense two-pass update vs the
    oracle's coalesce + row update, and vs the list-walk kernel (small_rows_max=0)."""
    from oracle import dlrm_numpy as O
    from dlrm_b200.engine import Engine

    rng = np.random.default_rng(9)
    D, ln_emb, B = 128, [3, 6000, 10, 155], 4096
    W = [rng.standard_normal((n, D)).astype(np.float32) for n in ln_emb]
    X, off, idx = O.random_batch(rng, ln_emb, B, 13, 3)
    off = [o.astype(idx_dtype) for o in off]
    idx = [i.astype(idx_dtype) for i in idx]
    dY = (rng.standard_normal((B, len(ln_emb) + 1, D)) * 1e-2).astype(np.float32)
    res = {}
    for small_max in (256, 0):
        e = Engine(D, ln_emb, [13, D], [D + 10, 1], device=DEV, max_batch=B, gemm="simt", small_rows_max=small_max)
        e.load_params(dict(emb=W, bot=[(np.zeros((D, 13), np.float32), np.zeros(D, np.float32))],
                           top=[(np.zeros((1, D + 10), np.float32), np.zeros(1, np.float32))]))
        e.ensure_optimizer_state(opt)
        assert [e.is_small(k) for k in range(4)] == ([True, False, True, True] if small_max else [False] * 4)
        sp = _sp(off, idx)
        e.dT.copy_(torch.from_numpy(dY))
        for _ in range(2):
            e.emb_link(sp)
            e.emb_update(sp, optimizer=opt, lr=0.05)
        torch.cuda.synchronize()
        res[small_max] = ([e.table(k).cpu().numpy() for k in range(4)],
                          e.momentum.cpu().numpy() if opt == "rwsadagrad" else None)
        assert int(e.head.abs().sum().item()) == 0
    # oracle: coalesced sparse update, twice
    Wr = [w.copy() for w in W]
    mom = [np.zeros(n, np.float32) for n in ln_emb]
    for _ in range(2):
        for k, n in enumerate(ln_emb):
            bag = np.searchsorted(off[k], np.arange(idx[k].size), side="right") - 1
            g = np.zeros((n, D), np.float64)
            np.add.at(g, idx[k].astype(np.int64), dY[bag, 1 + k, :].astype(np.float64))
            rows = np.unique(idx[k])
            gr = g[rows].astype(np.float32)
            if opt == "rwsadagrad":
                mom[k][rows] += (gr * gr).mean(axis=1)
                Wr[k][rows] -= 0.05 * gr / (np.sqrt(mom[k][rows]) + 1e-10)[:, None]
            else:
                Wr[k][rows] -= 0.05 * gr
    for small_max in (256, 0):
        for k in range(4):
            np.testing.assert_allclose(res[small_max][0][k], Wr[k], rtol=2e-4, atol=2e-5)
        if opt == "rwsadagrad":
            np.testing.assert_allclose(res[small_max][1], np.concatenate(mom), rtol=2e-4, atol=1e-9)


@pytest.mark.parametrize("idx_bytes", [4, 8])
def test_device_multihot_generator_is_bitwise_the_host_generator(idx_bytes):
    from dlrm_b200 import _lib, mlperf as M

    lib = _lib.lib()
    tables = [0, 5, 8, 11, 20, 21, 25]
    B, sample0, seed, step = 512, 8192 * 3, 7, 42
    dt = torch.int32 if idx_bytes == 4 else torch.int64
    outs = [torch.zeros(B * M.MULTI_HOT[t], dtype=dt, device=DEV) for t in tables]
    X = torch.zeros((B, 13), dtype=torch.float32, device=DEV)
    T = torch.zeros(B, dtype=torch.float32, device=DEV)
    n = len(tables)
    _lib.check(lib.dlrm_b200_gen_multihot((C.c_void_p * n)(*[o.data_ptr() for o in outs]),
                                          (C.c_int64 * n)(*[M.TABLE_ROWS[t] for t in tables]),
                                          (C.c_int * n)(*[M.MULTI_HOT[t] for t in tables]), (C.c_int * n)(*tables), n,
                                          idx_bytes, seed, step, sample0, B, X.data_ptr(), T.data_ptr(), 13, None),
               "gen_multihot")
    torch.cuda.synchronize()
    want = M.multi_hot_batch(seed, step, M.TABLE_ROWS, M.MULTI_HOT, sample0, B, tables=tables, dtype=np.int64)
    for o, w, t in zip(outs, want, tables):
        assert np.array_equal(o.cpu().numpy().astype(np.int64).reshape(B, -1), w), t
    Xw, Tw = M.dense_and_targets(seed, step, sample0, B, 13)
    assert np.array_equal(X.cpu().numpy(), Xw) and np.array_equal(T.cpu().numpy(), Tw[:, 0])


def test_index_outside_its_table_is_reported_not_dereferenced():
    """nn.EmbeddingBag raises on an out-of-range index (round-1 advisor finding: silent corruption here)."""
    from oracle import dlrm_numpy as O
    from dlrm_b200 import _lib
    from dlrm_b200.engine import Engine

    rng = np.random.default_rng(3)
    D, ln_emb, B = 32, [100, 50], 64
    e = Engine(D, ln_emb, [13, D], [D + 3, 1], device=DEV, max_batch=B, gemm="simt")
    e.init_params(0)
    guard = e.tables.clone()
    X, off, idx = O.random_batch(rng, ln_emb, B, 13, 4)
    sp = _sp(off, idx)
    e.emb_forward(sp)
    assert e.lib.dlrm_b200_check_device_errors(None) == 0
    idx[1][3] = 50                      # == rows of table 1
    idx[0][0] = -7
    sp = _sp(off, idx)
    e.ensure_optimizer_state("sgd")
    e.emb_forward(sp, link=True)
    e.emb_update(sp, optimizer="sgd", lr=0.0)
    torch.cuda.synchronize()
    assert e.lib.dlrm_b200_check_device_errors(None) != 0
    assert b"outside its table" in e.lib.dlrm_b200_last_error()
    assert e.lib.dlrm_b200_check_device_errors(None) == 0       # cleared by the read
    assert torch.equal(guard, e.tables)                          # lr = 0 and nothing written out of place


def test_block_copy_and_reduce_partials():
    from dlrm_b200 import _lib

    lib = _lib.lib()
    srcs = [torch.randn(n, device=DEV) for n in (4, 1024, 100000)]
    dsts = [torch.zeros_like(s) for s in srcs]
    n = len(srcs)
    _lib.check(lib.dlrm_b200_block_copy((C.c_void_p * n)(*[s.data_ptr() for s in srcs]),
                                        (C.c_void_p * n)(*[d.data_ptr() for d in dsts]),
                                        (C.c_int64 * n)(*[s.numel() * 4 for s in srcs]), n, None), "block_copy")
    for s, d in zip(srcs, dsts):
        assert torch.equal(s, d)
    B, F, D = 37, 6, 64
    part = torch.randn(5, B, D, device=DEV)          # slabs 0-1 -> feature 2, slabs 2-4 -> feature 5
    T = torch.zeros(B, F, D, device=DEV)
    _lib.check(lib.dlrm_b200_emb_reduce_partials(part.data_ptr(), T.data_ptr(), F * D, B, D, (C.c_int * 2)(2, 5),
                                                 (C.c_int * 3)(0, 2, 5), 2, None), "reduce_partials")
    assert torch.equal(T[:, 2], part[0] + part[1])
    assert torch.equal(T[:, 5], (part[2] + part[3]) + part[4])
    assert float(T[:, [0, 1, 3, 4]].abs().sum()) == 0.0
