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


# ---------------------------------------------------------------------------- routes of the peer-memory exchange
# Pure index arithmetic (offsets in ELEMENTS, not bytes), shared by the engine / DistEngine / MultiHotExchange and by
# the CPU emulation of an N-rank step in tests/test_shard_routes.py.

def out_routes(shards: Sequence[dict], slots, B: int, F: int, D: int):
    """Where a rank's gather puts the pooled rows of its local shards inside the TP buffer of the rank that owns the
    sample -- TP = T [B, F, D] followed by the partial-sum area [slab][B][D], one slab per (row-split table, part):
    route_out[j] = (offset of sample 0, sample stride); route_dy[j] = offset of shard j's gradient row inside one
    sample of dT [B, F, D].  `slots` = split_slots(placement): (table, nparts) of every row-split table."""
    first, slot_of, acc = [], {}, 0
    for i, (t, n) in enumerate(slots):
        slot_of[int(t)] = i
        first.append(acc)
        acc += int(n)
    route_out, route_dy = [], []
    for sh in shards:
        t = int(sh["table"])
        if int(sh["nparts"]) == 1:
            route_out.append(((1 + t) * D, F * D))
        else:
            slab = first[slot_of[t]] + int(sh["part"])
            route_out.append((B * F * D + slab * B * D, D))
        route_dy.append((1 + t) * D)
    return route_out, route_dy


def grad_routes(pl: Placement, rank: int, B: int, D: int, F: int):
    """Destinations of rank `rank`'s interaction backward: feature 0 stays in its own dT (dst rank -1); the gradient
    row of table t (feature 1 + t) goes to slab `rank` of the receive buffer [world][B][Tl_dst][D] of EVERY rank storing
    rows of t.  Returns ([(dst rank, element offset of local sample 0 inside that buffer, sample stride)], first) with
    the destinations of feature f at [first[f], first[f + 1])."""
    n_tables = 1 + max(s.table for s in pl.shards) if pl.shards else 0
    dst, first = [(-1, 0, F * D)], [0, 1]
    for t in range(n_tables):
        for s in pl.of_table(t):
            own = pl.of_rank(s.rank)
            j = own.index(s)
            dst.append((s.rank, ((rank * B) * len(own) + j) * D, len(own) * D))
        first.append(len(dst))
    return dst, first


def index_copies(pl: Placement, rank: int, hot: Sequence[int], B: int):
    """Fixed-length bags: the index block of table t that rank `rank` uploads ([B, L_t], its own samples) goes to slot
    `rank` of the global index array [world, B, L_t] of every rank storing rows of t.  Returns
    [(table, dst rank, local shard index on dst, element offset inside dst's array, elements)]."""
    out = []
    for t, L in enumerate(hot):
        for s in pl.of_table(t):
            j = pl.of_rank(s.rank).index(s)
            out.append((t, s.rank, j, rank * B * int(L), B * int(L)))
    return out
