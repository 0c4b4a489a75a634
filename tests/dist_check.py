"""Multi-GPU parity check of dlrm_b200.dist.DistEngine, launched with torchrun (1 process per GPU):

    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py

Every rank also builds the FULL model in a single-GPU Engine and runs the global batch through
it: the sharded forward must reproduce the slice of the single-device logits, and one SGD step
must leave the local tables / the MLPs where the single-device run puts them (embedding
gradients are summed over ranks, not averaged -- the reference's all-to-all semantics -- which
for SGD equals a single-device step with lr * world on the tables).  Test infrastructure."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dlrm_b200.dist import DistEngine, init_distributed, table_slices  # noqa: E402
from dlrm_b200.engine import Engine, sparse_from_reference  # noqa: E402
from oracle import dlrm_numpy as O  # noqa: E402  (checker only)


def main():
    rank, world = init_distributed("nccl")
    dev = "cuda:%d" % int(os.environ.get("LOCAL_RANK", rank))
    gemm = os.environ.get("DLRM_GEMM", "tc")
    exchange = os.environ.get("DLRM_EXCHANGE", "nccl")
    D, ln_emb, ln_bot = 128, [3000, 500, 40, 1000, 77], [13, 64, 128]
    if world > 4:      # uneven table-wise slices (2 or 1 tables per rank at world 8)
        ln_emb = ln_emb + [250, 1200, 64, 900, 333, 2100]
    Tg = len(ln_emb)
    ln_top = [D + (Tg + 1) * Tg // 2, 64, 32, 1]
    B = 96
    Bg = B * world
    rng = np.random.default_rng(5)
    params = O.random_params(rng, D, ln_emb, ln_bot, ln_top)
    X, off, idx = O.random_batch(rng, ln_emb, Bg, 13, 10)
    tgt = np.round(rng.random((Bg, 1))).astype(np.float32)
    # ---- single-device run of the global batch
    full = Engine(D, ln_emb, ln_bot, ln_top, loss="bce", sigmoid_top=len(ln_top) - 2, device=dev, max_batch=Bg,
                  gemm=gemm)
    full.load_params(params)
    spg = sparse_from_reference([torch.from_numpy(o) for o in off], [torch.from_numpy(i) for i in idx], dev)
    Xg, Tgt = torch.from_numpy(X).to(dev), torch.from_numpy(tgt).to(dev)
    p_full = full.forward(Xg, spg).clone()
    # ---- sharded run
    t0, t1 = table_slices(Tg, world)[rank]
    de = DistEngine(D, ln_emb, ln_bot, ln_top, local_batch=B, device=dev, gemm=gemm, exchange=exchange)
    loc = dict(emb=params["emb"][t0:t1], bot=params["bot"], top=params["top"], v_W_l=None)
    de.eng.load_params(loc)
    spl = sparse_from_reference([torch.from_numpy(o) for o in off[t0:t1]],
                                [torch.from_numpy(i) for i in idx[t0:t1]], dev)
    sl = slice(rank * B, (rank + 1) * B)
    Xl, Tl = Xg[sl].contiguous(), Tgt[sl].contiguous()
    p_loc = de.forward(Xl, spl)
    err_f = float((p_loc - p_full[sl]).abs().max().item())
    ok = err_f < 2e-6
    # ---- one SGD step
    lr = 0.05
    full.forward(Xg, spg, link=True, skip_head=True)
    full.backward(Xg, spg, Tgt)
    full.emb_update(spg, full.dT.view(-1)[D:], full.F * D, D, "sgd", lr * world)
    if full.tc:
        full._dense_update_pack(0, lr)
    else:
        full.dense_step("sgd", lr)
    loss_loc = de.train_step(Xl, spl, Tl, lr, "sgd")
    torch.cuda.synchronize()
    err_t = 0.0
    for j, k in enumerate(range(t0, t1)):
        err_t = max(err_t, float((de.eng.table(j) - full.table(k)).abs().max().item()))
    err_d = float((de.eng.dense - full.dense).abs().max().item())
    ok = ok and err_t < 2e-6 and err_d < 2e-6
    # global mean loss == mean of local losses
    ll = loss_loc.clone()
    dist.all_reduce(ll, op=dist.ReduceOp.AVG)
    err_l = abs(float(ll.item()) - float(full.loss_buf.item()))
    ok = ok and err_l < 2e-6
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    print("rank %d exchange=%s gemm=%s fwd_err=%.2e table_err=%.2e dense_err=%.2e loss_err=%.2e -> %s" % (
        rank, exchange, gemm, err_f, err_t, err_d, err_l, "PASS" if ok else "FAIL"), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
