"""Packed batch format and the bench's synthetic generator (dlrm_b200/data.py) -- CPU only."""
import numpy as np
import pytest
import torch

from dlrm_b200.data import HostBatch, PackedLayout, fill_batch, make_batch


@pytest.mark.parametrize("fixed,per_table_L", [(False, None), (True, None), (False, [3, 1, 7, 2])])
def test_generated_batch_obeys_the_format(fixed, per_table_L):
    ln_emb, B, lmax = [1000, 7, 100000, 2], 64, 5
    hb = make_batch(np.random.default_rng(3), ln_emb, B, m_den=13, lmax=lmax, fixed=fixed, per_table_L=per_table_L,
                    pin=False)
    L = hb.layout
    off = hb.offsets
    assert off.shape == (4, B + 1) and off[0, 0] == 0 and off[-1, -1] == hb.nnz
    assert (off[1:, 0] == off[:-1, -1]).all()                 # tables are contiguous in one position space
    assert (np.diff(off, axis=1) >= 0).all()
    idx = hb.indices
    for k, R in enumerate(ln_emb):
        ids = idx[off[k, 0]:off[k, B]]
        assert ids.min() >= 0 and ids.max() < R
        lens = np.diff(off[k])
        if per_table_L is not None:
            assert (lens == per_table_L[k]).all()             # multi-hot style: fixed length, repeats allowed
        else:
            assert lens.min() >= 1 and lens.max() <= min(R, lmax)
            for b in range(B):                                  # like np.unique in the reference generator
                bag = ids[off[k, b] - off[k, 0]:off[k, b + 1] - off[k, 0]]
                assert (np.diff(bag) > 0).all()
    assert hb.X.shape == (B, 13) and hb.X.dtype == torch.float32
    assert set(np.unique(hb.target.numpy())) <= {0.0, 1.0}
    assert hb.used_bytes == L.off_i + hb.nnz * 8 <= L.nbytes
    # reference view: per-table local start offsets [T, B] and per-table id tensors
    X, lS_o, lS_i, T = hb.reference_format()
    assert lS_o.shape == (4, B) and (lS_o[:, 0] == 0).all()
    assert sum(int(i.numel()) for i in lS_i) == hb.nnz
    for k in range(4):
        assert np.array_equal(lS_o[k].numpy(), off[k, :B] - off[k, 0])
        assert np.array_equal(lS_i[k].numpy(), idx[off[k, 0]:off[k, B]])


def test_capacity_is_checked_and_views_alias_one_buffer():
    hb = HostBatch(PackedLayout(8, 2, 4, 10), pin=False)
    with pytest.raises(RuntimeError):
        fill_batch(hb, np.random.default_rng(0), [1000, 1000], lmax=10, fixed=True)
    hb = make_batch(np.random.default_rng(0), [50, 60], 8, m_den=4, lmax=3, pin=False)
    base = hb.buf.data_ptr()
    L = hb.layout
    assert hb.X.data_ptr() == base + L.off_x and hb.target.data_ptr() == base + L.off_t
    assert hb.offsets_t.data_ptr() == base + L.off_o and hb.indices_t.data_ptr() == base + L.off_i
