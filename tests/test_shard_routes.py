"""CPU emulation of one sharded step's DATA MOVEMENT at world sizes 1, 2, 4 and 8 with the route arithmetic the
engine, DistEngine and MultiHotExchange use on the GPU (dlrm_b200/sharding.py: out_routes, grad_routes, index_copies).
Ranks are numpy buffers in one process; every store a kernel would make through peer-mapped memory is made here with
the same offsets.  Checked: every pooled row / partial sum lands exactly once in the right slot of the sample owner's
interaction operand, the partial sums of a row-split table add up, every gradient row reaches slab `src` of every
rank storing rows of its table at the place that rank's update reads, and the index blocks assemble the global
[world, B, L] arrays."""
import numpy as np
import pytest

from dlrm_b200 import mlperf as M, placement as P, sharding as S


def _f(t, b, d):            # pooled value of (table, global sample, column)
    return 1000.0 * t + 10.0 * b + d


def _placements():
    rows, cost = list(M.TABLE_ROWS), [float(h) for h in M.MULTI_HOT]
    out = []
    for W in (1, 2, 4, 8):
        out.append(("mlperf", W, P.plan(rows, cost, W), len(rows)))
        out.append(("three_tables", W, P.plan([1000, 1000, 1000], [5.0] * 3, W, force_split=[0] if W <= 2 else []), 3))
        out.append(("equal26", W, P.plan([1_000_000] * 26, [5.05] * 26, W), 26))
    return out


@pytest.mark.parametrize("name,W,pl,Tg", _placements(), ids=lambda v: str(v) if isinstance(v, (str, int)) else "")
def test_emulated_exchange(name, W, pl, Tg):
    B, D = 3, 4
    F = Tg + 1
    slots = S.split_slots(pl)
    n_slabs = sum(n for _, n in slots)
    hot = [1 + (t % 3) for t in range(Tg)]
    mine = [pl.of_rank(r) for r in range(W)]
    TP = [np.full(B * F * D + n_slabs * B * D, np.nan) for _ in range(W)]
    hits = [np.zeros(B * F * D + n_slabs * B * D, dtype=np.int64) for _ in range(W)]
    grecv = [np.full(W * B * max(len(mine[r]), 1) * D, np.nan) for r in range(W)]
    ghits = [np.zeros_like(g, dtype=np.int64) for g in grecv]
    idxg = [[np.full(W * B * hot[s.table], -1, dtype=np.int64) for s in mine[r]] for r in range(W)]

    # ---- forward: rank s pools its shards for the GLOBAL batch and stores into the sample owner's TP
    for s in range(W):
        ek = S.engine_kwargs(pl, s, Tg)
        route_out, _ = S.out_routes(ek["shards"], ek["split_slots"], B, F, D)
        for j, sh in enumerate(mine[s]):
            off, stride = route_out[j]
            for b in range(W * B):
                o, bl = divmod(b, B)
                for d in range(D):
                    a = off + bl * stride + d
                    TP[o][a] = _f(sh.table, b, d) * (sh.part + 1)
                    hits[o][a] += 1
    slab_first = np.concatenate([[0], np.cumsum([n for _, n in slots])]).astype(int)
    for o in range(W):
        T = TP[o][:B * F * D].reshape(B, F, D)
        part = TP[o][B * F * D:].reshape(max(n_slabs, 0), B, D) if n_slabs else None
        for i, (t, n) in enumerate(slots):                 # reduce_partials: fixed part order
            T[:, 1 + t, :] = part[slab_first[i]:slab_first[i] + n].sum(axis=0)
            hits[o][:B * F * D].reshape(B, F, D)[:, 1 + t, :] += 1
        assert (hits[o][:B * F * D].reshape(B, F, D)[:, 1:, :] == 1).all()      # every slot written exactly once
        assert (hits[o][:B * F * D].reshape(B, F, D)[:, 0, :] == 0).all()       # the dense feature is not the gather's
        assert (hits[o][B * F * D:] == 1).all()
        for t in range(Tg):
            n = len(pl.of_table(t))
            scale = n * (n + 1) / 2
            for bl in range(B):
                for d in range(D):
                    assert T[bl, 1 + t, d] == _f(t, o * B + bl, d) * scale

    # ---- backward: rank o's interaction backward pushes the gradient row of (sample, table) to every owner
    for o in range(W):
        routes, first = S.grad_routes(pl, o, B, D, F)
        assert routes[0][0] == -1 and first[:2] == [0, 1] and len(first) == F + 1
        for t in range(Tg):
            assert first[2 + t] - first[1 + t] == len(pl.of_table(t))
            for r, off, stride in routes[first[1 + t]:first[2 + t]]:
                for bl in range(B):
                    for d in range(D):
                        a = off + bl * stride + d
                        grecv[r][a] = -_f(t, o * B + bl, d)
                        ghits[r][a] += 1
    for r in range(W):
        Tl = len(mine[r])
        if Tl == 0:
            continue
        assert (ghits[r] == 1).all()
        slab, dy_stride = B * Tl * D, Tl * D               # DistEngine._setup_p2p: peer_dY[s] = base + s * slab
        for j, sh in enumerate(mine[r]):
            for b in range(W * B):
                src, bl = divmod(b, B)
                for d in range(D):
                    assert grecv[r][src * slab + bl * dy_stride + j * D + d] == -_f(sh.table, b, d)

    # ---- fixed-length bags: every rank uploads ITS samples; blocks assemble [world, B, L] at every owner
    for o in range(W):
        for t, r, j, off, n in S.index_copies(pl, o, hot, B):
            assert mine[r][j].table == t and n == B * hot[t]
            block = np.array([[1_000_000 * t + 100 * (o * B + bl) + l for l in range(hot[t])] for bl in range(B)])
            assert (idxg[r][j][off:off + n] == -1).all()
            idxg[r][j][off:off + n] = block.reshape(-1)
    for r in range(W):
        for j, sh in enumerate(mine[r]):
            L = hot[sh.table]
            got = idxg[r][j].reshape(W * B, L)
            for b in range(W * B):
                assert (got[b] == [1_000_000 * sh.table + 100 * b + l for l in range(L)]).all()
