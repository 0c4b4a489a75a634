"""MLPerf-DLRM (Criteo Terabyte, TorchRec variant) synthetic workload: BASELINE.json configs[3] ("CFG3").

  * table sizes           torchrec_dlrm/README.MD:45   (204,184,588 rows x dim 128 = 104.5 GB fp32)
  * multi-hot sizes L_k   torchrec_dlrm/README.MD:159  (sum 214 lookups per sample)
  * MLPs                  bench/run_and_time.sh:17     bot 13-512-256-128, top 1024-1024-512-256-1
  * bag construction      torchrec_dlrm/multi_hot.py:80-127: every table owns a FIXED lookup table
        multi_hot_table[r] = [r, u_1(r), ..., u_{L-1}(r)],   u_j(r) ~ U{0..R-1} drawn once (seed 0),
    and a sample whose 1-hot id is r gets the bag multi_hot_table[r] (fixed length L, duplicates allowed).

The reference materialises those lookup tables (40M x 100 int32 = 16 GB for one table).  Here the table is a
counter-based hash: u_j(r) = mulhi64(splitmix64(key(table, r, j)), R) -- the same object (a fixed function of
(table, row, slot), uniform over the rows) without the memory, identical on host (numpy, below) and device
(csrc/datagen.cu), so generated batches are bit-identical on both sides.  The 1-hot ids of a synthetic batch
are drawn the same way from (seed, step, table, sample).
"""
from __future__ import annotations

from typing import Sequence

import numpy as np

TABLE_ROWS = [40000000, 39060, 17295, 7424, 20265, 3, 7122, 1543, 63, 40000000, 3067956, 405282, 10, 2209,
              11938, 155, 4, 976, 14, 40000000, 40000000, 40000000, 590152, 12973, 108, 36]
MULTI_HOT = [3, 2, 1, 2, 6, 1, 1, 1, 1, 7, 3, 8, 1, 6, 9, 5, 1, 1, 1, 12, 100, 27, 10, 3, 1, 1]
DIM = 128
LN_BOT = [13, 512, 256, 128]
TOP_TAIL = [1024, 1024, 512, 256, 1]
GLOBAL_BATCH = 65536

_M64 = (1 << 64) - 1
_K_TABLE, _K_ROW, _K_SLOT, _K_STEP = 0x9E3779B97F4A7C15, 0xC2B2AE3D27D4EB4F, 0x165667B19E3779F9, 0xD6E8FEB86659FD93


def ln_top(n_tables: int = len(TABLE_ROWS), dim: int = DIM):
    return [dim + (n_tables + 1) * n_tables // 2] + TOP_TAIL


def splitmix64(x: np.ndarray) -> np.ndarray:
    """Finaliser of splitmix64 on uint64 arrays (wrap-around arithmetic)."""
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def mulhi64(x: np.ndarray, n: int) -> np.ndarray:
    """floor(x * n / 2^64) for uint64 x and 0 < n < 2^32: uniform map of a 64-bit hash onto [0, n)."""
    assert 0 < n < (1 << 32)
    x = np.asarray(x, dtype=np.uint64)
    n64 = np.uint64(n)
    lo = (x & np.uint64(0xFFFFFFFF)) * n64
    hi = (x >> np.uint64(32)) * n64
    return ((hi + (lo >> np.uint64(32))) >> np.uint64(32)).astype(np.int64)


def _key(a: int, b, c) -> np.ndarray:
    with np.errstate(over="ignore"):
        return (np.uint64(a) ^ (np.asarray(b, dtype=np.uint64) * np.uint64(_K_ROW))
                ^ (np.asarray(c, dtype=np.uint64) * np.uint64(_K_SLOT)))


def one_hot_ids(seed: int, step: int, table: int, rows: int, samples: np.ndarray) -> np.ndarray:
    """1-hot id of every (global) sample index in `samples` for `table` at `step`."""
    base = ((seed * _K_STEP) ^ ((step + 1) * _K_TABLE) ^ ((table + 1) * _K_SLOT)) & _M64
    return mulhi64(splitmix64(splitmix64(_key(base, samples, 0))), rows)


def multi_hot_row(table: int, rows: int, L: int, ids: np.ndarray) -> np.ndarray:
    """multi_hot_table[ids] -> [len(ids), L]: column 0 = the id itself, column j>0 = hash(table, id, j)."""
    ids = np.asarray(ids, dtype=np.int64)
    out = np.empty((ids.size, L), dtype=np.int64)
    out[:, 0] = ids
    if L > 1:
        base = (((table + 1) * _K_TABLE) ^ 0xA5A5A5A5A5A5A5A5) & _M64
        j = np.arange(1, L, dtype=np.uint64)[None, :]
        out[:, 1:] = mulhi64(splitmix64(_key(base, ids.astype(np.uint64)[:, None], j)), rows)
    return out


def multi_hot_batch(seed: int, step: int, rows: Sequence[int], hot: Sequence[int], sample0: int, batch: int,
                    tables=None, dtype=np.int32):
    """Indices of samples [sample0, sample0 + batch) of global step `step`: list (one per table in `tables`,
    default all) of [batch, L_k] arrays, row-major (bag b = row b: fixed-length bags, offsets = b * L_k)."""
    samples = np.arange(sample0, sample0 + batch, dtype=np.uint64)
    out = []
    for k in (range(len(rows)) if tables is None else tables):
        ids = one_hot_ids(seed, step, k, int(rows[k]), samples)
        out.append(multi_hot_row(k, int(rows[k]), int(hot[k]), ids).astype(dtype))
    return out


def dense_and_targets(seed: int, step: int, sample0: int, batch: int, m_den: int = 13):
    """Dense features U[0,1) fp32 [batch, m_den] and rounded targets [batch, 1] of the same samples (hashes of
    (seed, step, sample, column): every rank regenerates exactly its slice of the global batch)."""
    samples = np.arange(sample0, sample0 + batch, dtype=np.uint64)[:, None]
    cols = np.arange(m_den + 1, dtype=np.uint64)[None, :]
    base = ((seed * _K_STEP) ^ ((step + 1) * _K_TABLE) ^ 0x5DEECE66D) & _M64
    h = splitmix64(splitmix64(_key(base, samples, cols)))
    u = ((h >> np.uint64(40)).astype(np.float32)) * np.float32(1.0 / (1 << 24))
    return u[:, :m_den].copy(), np.round(u[:, m_den:]).astype(np.float32)
