"""Host-side logic of the sharded placement: planner, shard bookkeeping, multi-hot generator (CPU)."""
import numpy as np

from dlrm_b200 import mlperf as M, placement as P, sharding as S


def _check_partition(pl, rows):
    for t, R in enumerate(rows):
        sh = pl.of_table(t)
        assert sh[0].row_lo == 0 and sh[-1].row_hi == R
        for a, b in zip(sh, sh[1:]):
            assert a.row_hi == b.row_lo
        assert [s.part for s in sh] == list(range(len(sh))) and all(s.nparts == len(sh) for s in sh)


def test_mlperf_placement_is_balanced_and_fits():
    for w in (1, 2, 4, 8):
        pl = P.plan(M.TABLE_ROWS, M.MULTI_HOT, w)
        _check_partition(pl, M.TABLE_ROWS)
        assert pl.imbalance() < 1.06, (w, pl.load())
        assert max(pl.rows_per_rank()) * 512 < 150e9
    pl8 = P.plan(M.TABLE_ROWS, M.MULTI_HOT, 8)
    assert pl8.split_tables() == [20, 21]          # the L = 100 and L = 27 tables
    ref = P.contiguous(M.TABLE_ROWS, 8)
    ref.cost = [float(c) for c in M.MULTI_HOT]
    assert ref.imbalance() > 5.0                     # the reference's slices: one rank does 64 % of the lookups


def test_equal_tables_get_extra_splits():
    pl = P.plan([1_000_000] * 26, [5.05] * 26, 8)
    _check_partition(pl, [1_000_000] * 26)
    assert len(pl.split_tables()) == 2 and abs(pl.imbalance() - 1.0) < 1e-9
    assert P.plan([1000] * 3, [1.0] * 3, 8).imbalance() < 1.2     # fewer tables than ranks


def test_sharding_helpers():
    rows = [100, 7, 5000, 64]
    pl = P.plan(rows, [1, 1, 50, 1], 4)
    assert 2 in pl.split_tables()
    params = dict(emb=[np.arange(r * 2, dtype=np.float32).reshape(r, 2) for r in rows], bot=[], top=[], v_W_l=None)
    seen = {t: np.zeros(r, bool) for t, r in enumerate(rows)}
    for r in range(4):
        loc = S.slice_params(params, pl, r)
        kw = S.engine_kwargs(pl, r, len(rows))
        assert kw["ln_emb"] == [w.shape[0] for w in loc["emb"]] and kw["n_features"] == 5
        for sh, w in zip(kw["shards"], loc["emb"]):
            assert np.array_equal(w, params["emb"][sh["table"]][sh["row_lo"]:sh["row_lo"] + sh["row_n"]])
            assert not seen[sh["table"]][sh["row_lo"]:sh["row_lo"] + sh["row_n"]].any()
            seen[sh["table"]][sh["row_lo"]:sh["row_lo"] + sh["row_n"]] = True
        assert kw["split_slots"] == S.split_slots(pl)
    assert all(v.all() for v in seen.values())
    own = S.table_owners(pl)
    assert own[2] == [0, 1, 2, 3] and all(len(own[t]) == 1 for t in (0, 1, 3))


def test_multihot_generator_properties():
    b = M.multi_hot_batch(3, 5, M.TABLE_ROWS, M.MULTI_HOT, 1000, 256)
    for k, x in enumerate(b):
        assert x.shape == (256, M.MULTI_HOT[k]) and x.min() >= 0 and x.max() < M.TABLE_ROWS[k]
    # a rank's slice == the same samples of the global batch; the lookup rows are a fixed function of the 1-hot id
    g = M.multi_hot_batch(3, 5, M.TABLE_ROWS, M.MULTI_HOT, 0, 2048, tables=[9, 20])
    s = M.multi_hot_batch(3, 5, M.TABLE_ROWS, M.MULTI_HOT, 1024, 512, tables=[9, 20])
    assert all(np.array_equal(a[1024:1536], c) for a, c in zip(g, s))
    r1 = M.multi_hot_row(20, M.TABLE_ROWS[20], 100, g[1][:5, 0])
    assert np.array_equal(r1, g[1][:5].astype(np.int64))
    X, T = M.dense_and_targets(3, 5, 0, 4096)
    assert X.dtype == np.float32 and 0.0 <= X.min() and X.max() < 1.0 and abs(X.mean() - 0.5) < 0.01
    assert set(np.unique(T)) <= {0.0, 1.0} and abs(T.mean() - 0.5) < 0.05
    xs = np.array([0, 1, 2 ** 63, 2 ** 64 - 1, 123456789123456789], dtype=np.uint64)
    for n in (3, 40000000, 2 ** 32 - 1):
        assert list(M.mulhi64(xs, n)) == [(int(v) * n) >> 64 for v in xs]


def test_placement_properties_random_models():
    """For random table sets and world sizes: the shards of a table partition its rows exactly (no gap, no overlap,
    no empty part, parts numbered in row order with equal ceil-sized ranges -- what the remote-read kernel's
    `row / rows_per_shard` relies on), every table is placed, the plan is deterministic, and it is never worse
    balanced than the reference's contiguous slices."""
    from hypothesis import given, settings, strategies as st

    from dlrm_b200 import placement as P

    @settings(max_examples=200, deadline=None, derandomize=True, database=None)
    @given(st.lists(st.tuples(st.integers(1, 5_000_000), st.floats(0.5, 120.0)), min_size=1, max_size=30),
           st.sampled_from([1, 2, 3, 4, 8]))
    def check(tabs, world):
        rows, cost = [r for r, _ in tabs], [c for _, c in tabs]
        pl = P.plan(rows, cost, world)
        again = P.plan(rows, cost, world)
        assert [(s.table, s.rank, s.row_lo, s.row_hi, s.part, s.nparts) for s in pl.shards] == \
               [(s.table, s.rank, s.row_lo, s.row_hi, s.part, s.nparts) for s in again.shards]
        for t, n in enumerate(rows):
            sh = pl.of_table(t)
            assert sh and [s.part for s in sh] == list(range(len(sh))) and all(s.nparts == len(sh) and s.rows == n for s in sh)
            assert sh[0].row_lo == 0 and sh[-1].row_hi == n
            assert all(a.row_hi == b.row_lo for a, b in zip(sh, sh[1:])) and all(s.local_rows > 0 for s in sh)
            if len(sh) > 1:
                per = -(-n // len(sh))
                assert len(sh) == world and all(s.row_lo == s.part * per and s.rank == s.part for s in sh)
        assert abs(sum(pl.load()) - sum(cost)) < 1e-6 * sum(cost)
        if len(rows) >= world:
            ref = P.contiguous(rows, world)
            ref.cost = cost
            assert pl.imbalance() <= ref.imbalance() + 1e-3     # (ties within 1e-3 go to the heuristic)

    check()
