"""Input side of the hot path: synthetic batches in the reference's distribution and the packed
pinned-host / static-device batch buffers that replace dlrm_wrap's 28 pageable H2D copies
(dlrm_s_pytorch.py:129-145).

Distribution of `--data-generation=random` (dlrm_data_pytorch.py:899-960): dense X ~ U[0,1)
fp32 [B, m_den]; per (table, sample) a bag of round(max(1, u * min(R, Lmax))) draws
round(u * (R-1)), de-duplicated and sorted (np.unique); int64.  `per_table_L` gives the fixed
multi-hot sizes of the MLPerf/TorchRec variant (torchrec_dlrm/multi_hot.py:86-108).  The
generator is vectorised with its own RNG stream (the reference's is a Python double loop taking
0.8 s per 2048x26 batch); bit-identical reference inputs are only needed for parity and come from
tests/golden/.

Packed layout (one buffer, one H2D copy):
    [ X  f32 [B, m_den] | target f32 [B] | offsets i64 [T, B+1] | indices i64 [nnz] ]
offsets hold GLOBAL positions into `indices` (offsets[k, B] == offsets[k+1, 0]), so every table
shares one index pointer and all sizes live on the device: the kernels' `include_last` mode.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

from .engine import SparseInput


def _align(n: int, a: int = 16) -> int:
    return (n + a - 1) // a * a


@dataclass
class PackedLayout:
    B: int
    T: int
    m_den: int
    cap_nnz: int

    @property
    def off_x(self):
        return 0

    @property
    def off_t(self):
        return _align(self.B * self.m_den * 4)

    @property
    def off_o(self):
        return self.off_t + _align(self.B * 4)

    @property
    def off_i(self):
        return self.off_o + _align(self.T * (self.B + 1) * 8)

    @property
    def nbytes(self):
        return self.off_i + self.cap_nnz * 8

    def used(self, nnz: int) -> int:
        return self.off_i + nnz * 8


class HostBatch:
    """One batch in pinned host memory, packed."""

    def __init__(self, layout: PackedLayout, pin: bool = True):
        self.layout = layout
        self.buf = torch.empty(layout.nbytes, dtype=torch.uint8)
        if pin and torch.cuda.is_available():
            self.buf = self.buf.pin_memory()
        L = layout
        self.X = self.buf[L.off_x:L.off_x + L.B * L.m_den * 4].view(torch.float32).view(L.B, L.m_den)
        self.target = self.buf[L.off_t:L.off_t + L.B * 4].view(torch.float32).view(L.B, 1)
        self.offsets_t = self.buf[L.off_o:L.off_o + L.T * (L.B + 1) * 8].view(torch.int64).view(L.T, L.B + 1)
        self.indices_t = self.buf[L.off_i:L.off_i + L.cap_nnz * 8].view(torch.int64)
        self.nnz = 0

    # numpy views (zero copy)
    @property
    def offsets(self) -> np.ndarray:
        return self.offsets_t.numpy()

    @property
    def indices(self) -> np.ndarray:
        return self.indices_t.numpy()[:self.nnz]

    @property
    def used_bytes(self) -> int:
        return self.layout.used(self.nnz)

    def reference_format(self):
        """(X, lS_o [T,B], lS_i list, target) as the reference's loader yields them."""
        L = self.layout
        off = self.offsets
        lS_o = torch.from_numpy((off[:, :L.B] - off[:, :1]).copy())
        lS_i = [torch.from_numpy(self.indices[off[k, 0]:off[k, L.B]].copy()) for k in range(L.T)]
        return self.X.clone(), lS_o, lS_i, self.target.clone()


def fill_batch(hb: HostBatch, rng: Optional[np.random.Generator], ln_emb: Sequence[int], lmax: int = 10,
               fixed: bool = False, per_table_L: Optional[Sequence[int]] = None,
               round_targets: bool = True, table_seeds=None) -> HostBatch:
    """table_seeds: one seed (anything np.random.default_rng accepts) per table -> every table draws from its own
    stream, so that all ranks storing rows of a table see the same indices; X / target are then left alone."""
    L = hb.layout
    B = L.B
    if table_seeds is None:
        hb.X.numpy()[...] = rng.random((B, L.m_den), dtype=np.float32)
        t = rng.random((B, 1), dtype=np.float32)
        hb.target.numpy()[...] = np.round(t) if round_targets else t
    off = hb.offsets
    idx_all = hb.indices_t.numpy()
    pos = 0
    for k, R in enumerate(ln_emb):
        R = int(R)
        if table_seeds is not None:
            rng = np.random.default_rng(table_seeds[k])
        if per_table_L is not None:
            lens = np.full(B, int(per_table_L[k]), dtype=np.int64)
        elif fixed:
            lens = np.full(B, int(lmax), dtype=np.int64)
        else:
            lens = np.round(np.maximum(1.0, rng.random(B) * min(R, lmax))).astype(np.int64)
        tot = int(lens.sum())
        raw = np.round(rng.random(tot) * (R - 1)).astype(np.int64)
        bag = np.repeat(np.arange(B, dtype=np.int64), lens)
        if per_table_L is None:  # per-bag sorted unique (np.unique in the reference generator)
            order = np.lexsort((raw, bag))
            raw, bag = raw[order], bag[order]
            keep = np.ones(tot, dtype=bool)
            keep[1:] = (raw[1:] != raw[:-1]) | (bag[1:] != bag[:-1])
            raw, bag = raw[keep], bag[keep]
        n = raw.size
        if pos + n > L.cap_nnz:
            raise RuntimeError("packed batch capacity %d exceeded" % L.cap_nnz)
        cnt = np.bincount(bag, minlength=B)
        off[k, 0] = pos
        off[k, 1:] = pos + np.cumsum(cnt)
        idx_all[pos:pos + n] = raw
        pos += n
    hb.nnz = pos
    return hb


def make_batch(rng, ln_emb, B, m_den=13, lmax=10, fixed=False, per_table_L=None, pin=True) -> HostBatch:
    if per_table_L is not None:
        cap = int(B * sum(int(l) for l in per_table_L))
    else:
        cap = int(B * sum(min(int(r), lmax) for r in ln_emb))
    hb = HostBatch(PackedLayout(B, len(ln_emb), m_den, cap), pin)
    return fill_batch(hb, rng, ln_emb, lmax, fixed, per_table_L)


class DeviceBatch:
    """Static device-side twin of a HostBatch (fixed addresses -> CUDA-graph replayable)."""

    def __init__(self, layout: PackedLayout, device):
        self.layout = layout
        L = layout
        self.buf = torch.zeros(L.nbytes, dtype=torch.uint8, device=device)
        self.X = self.buf[L.off_x:L.off_x + L.B * L.m_den * 4].view(torch.float32).view(L.B, L.m_den)
        self.target = self.buf[L.off_t:L.off_t + L.B * 4].view(torch.float32).view(L.B, 1)
        self.offsets = self.buf[L.off_o:L.off_o + L.T * (L.B + 1) * 8].view(torch.int64).view(L.T, L.B + 1)
        self.indices = self.buf[L.off_i:L.off_i + L.cap_nnz * 8].view(torch.int64)
        self.sparse = SparseInput([self.indices] * L.T, [self.offsets[k] for k in range(L.T)], L.B, True,
                                  L.cap_nnz)
        self.nnz = 0

    def load(self, hb: HostBatch, non_blocking: bool = True):
        """ONE host->device copy of the used prefix of the packed buffer."""
        n = hb.used_bytes
        self.buf[:n].copy_(hb.buf[:n], non_blocking=non_blocking)
        self.nnz = hb.nnz
        return n


def to_device_packed(hb: HostBatch, device) -> DeviceBatch:
    db = DeviceBatch(hb.layout, device)
    db.load(hb, non_blocking=False)
    return db
