"""MLPerf binary record reader at the boundary of the hot path.

The reference's `CriteoBinDataset` (data_loader_terabyte.py:197-249) reads one mini-batch per item from
a flat file of int32 records `[label | 13 dense counts | 26 categorical ids]`, then `_transform_features`
(:74-93) turns it into the model's inputs: dense = log(x + 1) as fp32, ids optionally folded by
`max_ind_range`, one id per (table, sample) so the offsets of every table are 0..B-1.

This module keeps that surface (`CriteoBinDataset(data_file, counts_file, batch_size, max_ind_range,
bytes_per_feature)`, `len()`, `[idx]` -> `(x_int, lS_o, x_cat.t(), y)`) over a read-only memory map,
and adds `fill(idx, host_batch)`, which writes the batch straight into the packed pinned buffer of
dlrm_b200/data.py (one host-to-device copy per step).  `numpy_to_binary` writes such a file from
arrays (the train-split branch of data_loader_terabyte.py:252-290).
"""
from __future__ import annotations

import math
import os

import numpy as np
import torch

TAR_FEA, DEN_FEA, SPA_FEA = 1, 13, 26
TOT_FEA = TAR_FEA + DEN_FEA + SPA_FEA


class CriteoBinDataset(torch.utils.data.Dataset):
    """One item = one mini-batch of `batch_size` records (the last one may be short)."""

    def __init__(self, data_file, counts_file=None, batch_size=1, max_ind_range=-1, bytes_per_feature=4):
        if bytes_per_feature != 4:
            raise ValueError("records are int32: bytes_per_feature must be 4")
        self.tar_fea, self.den_fea, self.spa_fea = TAR_FEA, DEN_FEA, SPA_FEA
        self.tad_fea = TAR_FEA + DEN_FEA
        self.tot_fea = TOT_FEA
        self.batch_size = int(batch_size)
        self.max_ind_range = max_ind_range
        self.bytes_per_entry = bytes_per_feature * TOT_FEA * self.batch_size
        nbytes = os.path.getsize(data_file)
        self.num_entries = math.ceil(nbytes / self.bytes_per_entry)
        self.num_records = nbytes // (bytes_per_feature * TOT_FEA)
        print("data file:", data_file, "number of batches:", self.num_entries)
        self.records = np.memmap(data_file, dtype=np.int32, mode="r",
                                 shape=(self.num_records, TOT_FEA)) if self.num_records else \
            np.zeros((0, TOT_FEA), dtype=np.int32)
        self.counts = None
        if counts_file is not None:
            with np.load(counts_file) as data:
                self.counts = data["counts"]
        self.m_den = DEN_FEA

    def __len__(self):
        return self.num_entries

    def _rows(self, idx):
        lo = idx * self.batch_size
        return self.records[lo:min(lo + self.batch_size, self.num_records)]

    def __getitem__(self, idx):
        rec = torch.from_numpy(np.array(self._rows(idx)))          # private copy, like file.read()
        cat = rec[:, self.tad_fea:]
        if self.max_ind_range > 0:
            cat = cat % self.max_ind_range
        x_int = torch.log(rec[:, TAR_FEA:self.tad_fea].to(torch.float) + 1)
        y = rec[:, 0].to(torch.float32).view(-1, 1)
        n = rec.shape[0]
        lS_o = torch.arange(n).reshape(1, -1).repeat(SPA_FEA, 1)
        return x_int, lS_o, cat.to(torch.long).t(), y

    def fill(self, idx, hb):
        """Batch `idx` into a packed HostBatch: X = log(dense + 1), target, offsets [26, B+1] as global
        positions (table k owns positions k*B .. k*B+B), ids table-major."""
        rec = self._rows(idx)
        L = hb.layout
        n = rec.shape[0]
        if n != L.B or L.T != SPA_FEA or L.m_den != DEN_FEA:
            raise ValueError("record batch (%d x 13 dense x 26 ids) does not match the packed layout "
                             "(B=%d, T=%d, m_den=%d)" % (n, L.B, L.T, L.m_den))
        if n * SPA_FEA > L.cap_nnz:
            raise RuntimeError("packed batch capacity %d exceeded" % L.cap_nnz)
        dense = torch.from_numpy(np.ascontiguousarray(rec[:, TAR_FEA:TAR_FEA + DEN_FEA]))
        torch.log(dense.to(torch.float) + 1, out=hb.X)
        hb.target.numpy()[:, 0] = rec[:, 0]
        cat = rec[:, TAR_FEA + DEN_FEA:].astype(np.int64)
        if self.max_ind_range > 0:
            cat = cat % self.max_ind_range
        hb.indices_t.numpy()[:n * SPA_FEA].reshape(SPA_FEA, n)[...] = cat.T
        hb.offsets[...] = (np.arange(SPA_FEA, dtype=np.int64) * n)[:, None] + np.arange(n + 1, dtype=np.int64)[None, :]
        hb.nnz = n * SPA_FEA
        return hb


def numpy_to_binary(y, x_int, x_cat, output_file_path):
    """Write records `[y | x_int | x_cat]` as int32 (all values must fit into int32)."""
    y = np.asarray(y).reshape(-1, 1)
    rec = np.concatenate([y, np.asarray(x_int), np.asarray(x_cat)], axis=1).astype(np.int32)
    if rec.shape[1] != TOT_FEA:
        raise ValueError("expected 1 + 13 + 26 columns, got %d" % rec.shape[1])
    with open(output_file_path, "wb") as f:
        f.write(rec.tobytes())
