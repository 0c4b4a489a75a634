"""MLPerf binary record reader (dlrm_b200/binrecords.py) against the live reference's CriteoBinDataset
(when /root/reference is present) and against its documented transform; CPU only, exact equality."""
import os
import sys

import numpy as np
import pytest
import torch

from dlrm_b200 import binrecords as BR
from dlrm_b200.data import HostBatch, PackedLayout

REF = os.environ.get("DLRM_REFERENCE", "/root/reference")


def _make_file(tmp_path, n, seed=0):
    rng = np.random.RandomState(seed)
    y = rng.randint(0, 2, n)
    x_int = rng.randint(0, 5000, (n, 13))
    x_int[rng.rand(n, 13) < 0.2] = 0
    x_cat = rng.randint(0, 40_000_000, (n, 26))
    p = str(tmp_path / "recs.bin")
    BR.numpy_to_binary(y, x_int, x_cat, p)
    assert os.path.getsize(p) == n * 40 * 4
    return p, y, x_int, x_cat


@pytest.mark.parametrize("max_ind_range", [-1, 1000])
def test_items_follow_the_documented_transform_and_fill_matches(tmp_path, max_ind_range):
    n, B = 150, 64
    p, y, x_int, x_cat = _make_file(tmp_path, n)
    ds = BR.CriteoBinDataset(p, None, batch_size=B, max_ind_range=max_ind_range)
    assert len(ds) == 3
    for j in range(len(ds)):
        lo, hi = j * B, min((j + 1) * B, n)
        X, lS_o, lS_i, T = ds[j]
        cat = x_cat[lo:hi] % max_ind_range if max_ind_range > 0 else x_cat[lo:hi]
        assert torch.equal(X, torch.log(torch.tensor(x_int[lo:hi], dtype=torch.float) + 1))
        assert torch.equal(lS_i, torch.tensor(cat, dtype=torch.long).t()) and lS_i.shape == (26, hi - lo)
        assert torch.equal(lS_o, torch.arange(hi - lo).reshape(1, -1).repeat(26, 1))
        assert torch.equal(T, torch.tensor(y[lo:hi], dtype=torch.float32).view(-1, 1))
        hb = HostBatch(PackedLayout(hi - lo, 26, 13, (hi - lo) * 26), pin=False)
        ds.fill(j, hb)
        X2, o2, i2, T2 = hb.reference_format()
        assert torch.equal(X2, X) and torch.equal(T2, T) and torch.equal(o2, lS_o)
        assert torch.equal(torch.stack(i2), lS_i)
    with pytest.raises(ValueError):
        ds.fill(0, HostBatch(PackedLayout(B, 3, 13, B * 3), pin=False))


@pytest.mark.skipif(not os.path.isdir(REF), reason="live reference not present (GPU box)")
@pytest.mark.parametrize("max_ind_range", [-1, 77])
def test_against_live_reference_dataset(tmp_path, max_ind_range):
    sys.path.insert(0, REF)
    import builtins

    keep = builtins.print
    try:
        import data_loader_terabyte as dlt
    finally:
        builtins.print = keep
    n, B = 200, 32
    p, *_ = _make_file(tmp_path, n, seed=3)
    counts = str(tmp_path / "counts.npz")
    np.savez(counts, counts=np.arange(26))
    ref = dlt.CriteoBinDataset(p, counts, batch_size=B, max_ind_range=max_ind_range)
    mine = BR.CriteoBinDataset(p, counts, batch_size=B, max_ind_range=max_ind_range)
    assert len(ref) == len(mine) and np.array_equal(mine.counts, ref.counts)
    for j in range(len(ref)):
        for a, b in zip(ref[j], mine[j]):
            assert a.dtype == b.dtype and torch.equal(a, b)
