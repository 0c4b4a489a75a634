"""Property tests (hypothesis) of the host-side layout logic: packed batch layout, table slices,
exchange splits and the peer-store routes.  CPU only."""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

from dlrm_b200.data import PackedLayout
from dlrm_b200.dist import a2a_splits, push_route, table_slices


@settings(max_examples=200, deadline=None, derandomize=True, database=None)
@given(st.integers(1, 4096), st.integers(0, 64), st.integers(1, 64), st.integers(0, 100000))
def test_packed_layout_sections_are_aligned_and_disjoint(B, T, m_den, cap):
    L = PackedLayout(B, T, m_den, cap)
    sections = [(L.off_x, B * m_den * 4), (L.off_t, B * 4), (L.off_o, T * (B + 1) * 8), (L.off_i, cap * 8)]
    end = 0
    for off, n in sections:
        assert off % 16 == 0 and off >= end
        end = off + n
    assert L.nbytes >= end
    assert L.used(0) == L.off_i and L.used(cap) == L.off_i + cap * 8


@settings(max_examples=200, deadline=None, derandomize=True, database=None)
@given(st.integers(0, 300), st.integers(1, 16))
def test_table_slices_partition_like_the_reference_rule(n_tables, world):
    sl = table_slices(n_tables, world)
    assert len(sl) == world and sl[0][0] == 0 and sl[-1][1] == n_tables
    sizes = [e - s for s, e in sl]
    assert all(a[1] == b[0] for a, b in zip(sl, sl[1:]))
    assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)   # extras go to the first ranks


@settings(max_examples=100, deadline=None, derandomize=True, database=None)
@given(st.integers(1, 40), st.integers(1, 8), st.integers(1, 5), st.sampled_from([1, 2, 4, 16]))
def test_exchange_splits_and_push_routes_tile_the_receive_buffers(n_tables, world, B, D):
    sl = table_slices(n_tables, world)
    for r in range(world):
        send, recv = a2a_splits(n_tables, world, r, B, D)
        assert sum(recv) == B * n_tables * D and len(send) == len(recv) == world
        assert send == [B * (sl[r][1] - sl[r][0]) * D] * world
    # every element of every owner's receive buffer [world, B, T_r, D] is written exactly once
    hits = [np.zeros(world * B * (e - s) * D, dtype=np.int32) for s, e in sl]
    for src in range(world):
        route = push_route(n_tables, world, src, B, D)
        assert [o for o, _, _ in route] == [r for r, (s, e) in enumerate(sl) for _ in range(s, e)]
        for owner, off, ld in route:
            for b in range(B):
                hits[owner][off + b * ld: off + b * ld + D] += 1
    for h in hits:
        assert (h == 1).all()
