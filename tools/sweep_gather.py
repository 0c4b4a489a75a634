"""Tuning sweep of the multi-table gather (and the fused update) on the CFG1 workload.
    python tools/sweep_gather.py [out.json]
Prints achieved algorithmic GB/s per (bags_per_group, unroll) setting."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dlrm_b200 import _lib  # noqa: E402
from dlrm_b200.data import DeviceBatch, make_batch  # noqa: E402
from dlrm_b200.engine import Engine  # noqa: E402


def main():
    dev = "cuda:0"
    T, R, D, B = 26, 1_000_000, 128, 2048
    eng = Engine(D, [R] * T, [13, 512, 256, D], [D + (T + 1) * T // 2, 1024, 512, 256, 1], device=dev,
                 max_batch=B)
    eng.init_params(0)
    eng.ensure_optimizer_state("rwsadagrad")
    rng = np.random.default_rng(7)
    ring = []
    for _ in range(16):
        hb = make_batch(rng, [R] * T, B, 13, 10)
        db = DeviceBatch(hb.layout, dev)
        db.load(hb, non_blocking=False)
        ring.append(db)
    FD = eng.F * D
    out = eng.Tbuf.view(-1)[D:]
    res = []

    def timeit(fn, n=48):
        for i in range(6):
            fn(ring[i % 16])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(ring[i % 16])
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e-3

    nnz = float(np.mean([d.nnz for d in ring]))
    bytes_fwd = nnz * D * 4 + nnz * 8 + T * B * 8 + T * B * D * 4
    for S in (1, 2, 3, 4, 6, 8, 12, 16):
        for U in (4, 8):
            _lib.set_tunable("emb_bags_per_group", S)
            _lib.set_tunable("emb_unroll", U)
            t = timeit(lambda db: eng.emb_forward(db.sparse, out, FD, D))
            res.append(dict(kernel="gather", S=S, U=U, us=t * 1e6, GBs=bytes_fwd / t / 1e9))
            print(res[-1], flush=True)
    _lib.set_tunable("emb_bags_per_group", 0)
    _lib.set_tunable("emb_unroll", 0)
    # backward: link + update (RWSAdagrad); bytes per SURVEY §8(d)
    U_rows = nnz  # ~unique
    bytes_bwd = U_rows * (D * 4 * 2 + 8) + T * B * D * 4 + nnz * 8
    eng.dT.normal_()
    t_link = timeit(lambda db: eng.emb_link(db.sparse))
    eng.head.zero_()

    def both(db):
        eng.emb_link(db.sparse)
        eng.emb_update(db.sparse, eng.dT.view(-1)[D:], FD, D, "rwsadagrad", 0.01)

    t_both = timeit(both)
    res.append(dict(kernel="link", us=t_link * 1e6))
    res.append(dict(kernel="link+update", us=t_both * 1e6, update_us=(t_both - t_link) * 1e6,
                    update_GBs=bytes_bwd / (t_both - t_link) / 1e9))
    print(res[-2], res[-1], flush=True)

    # what bounds the update?  (a) no momentum traffic (SGD), (b) fused gather+link before it (list heads
    # evicted from L2 by the 140 MB of rows the gather streams through)
    def both_sgd(db):
        eng.emb_link(db.sparse)
        eng.emb_update(db.sparse, eng.dT.view(-1)[D:], FD, D, "sgd", 1e-6)

    t_sgd = timeit(both_sgd)
    res.append(dict(kernel="link+update(sgd)", us=t_sgd * 1e6, update_us=(t_sgd - t_link) * 1e6))
    print(res[-1], flush=True)

    def fused(db):
        eng.emb_forward(db.sparse, out, FD, D, link=True)
        eng.emb_update(db.sparse, eng.dT.view(-1)[D:], FD, D, "rwsadagrad", 1e-6)

    t_g = timeit(lambda db: eng.emb_forward(db.sparse, out, FD, D))
    t_f = timeit(fused)
    res.append(dict(kernel="gather+link -> update(rwsadagrad)", us=t_f * 1e6, gather_only_us=t_g * 1e6))
    print(res[-1], flush=True)
    if len(sys.argv) > 1:
        json.dump(res, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
