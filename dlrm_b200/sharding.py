"""Glue between a `placement.Placement` and the per-GPU `Engine`: which shards a rank stores, how the global
parameters / index streams map onto them.  Host logic only (no CUDA): unit-tested on CPU."""
from __future__ import annotations

from typing import Dict, List, Sequence

from .placement import Placement


def local_shards(pl: Placement, rank: int) -> List[dict]:
    """Shard descriptors of `rank` in the order of its local table arena (Engine(shards=...))."""
    return [dict(table=s.table, rows=s.rows, row_lo=s.row_lo, row_n=s.local_rows, part=s.part, nparts=s.nparts)
            for s in pl.of_rank(rank)]


def split_slots(pl: Placement):
    """(table, nparts) of every row-split table of the GLOBAL placement, ordered by table."""
    out = []
    for t in pl.split_tables():
        out.append((t, len(pl.of_table(t))))
    return out


def engine_kwargs(pl: Placement, rank: int, n_tables: int) -> dict:
    sh = local_shards(pl, rank)
    return dict(ln_emb=[s["row_n"] for s in sh], shards=sh, split_slots=split_slots(pl), n_features=n_tables + 1)


def slice_params(params: dict, pl: Placement, rank: int) -> dict:
    """Global parameter dict (oracle / state_dict layout: emb=[W_t], bot, top) -> the rank's view: one row slice
    per local shard."""
    out = dict(params)
    out["emb"] = [params["emb"][s.table][s.row_lo:s.row_hi] for s in pl.of_rank(rank)]
    if params.get("v_W_l") is not None:
        out["v_W_l"] = [params["v_W_l"][s.table][s.row_lo:s.row_hi] for s in pl.of_rank(rank)]
    return out


def local_streams(per_table: Sequence, pl: Placement, rank: int) -> List:
    """per_table[t] = anything describing table t's index stream (e.g. (indices, offsets) tensors of the global
    batch) -> the list ordered like the rank's shards (a table with several local shards appears several times)."""
    return [per_table[s.table] for s in pl.of_rank(rank)]


def table_owners(pl: Placement) -> Dict[int, List[int]]:
    """table -> ranks that need its indices / gradient rows (one for a whole table, all for a row-split one)."""
    out: Dict[int, List[int]] = {}
    for s in pl.shards:
        out.setdefault(s.table, [])
        if s.rank not in out[s.table]:
            out[s.table].append(s.rank)
    return {t: sorted(r) for t, r in out.items()}
