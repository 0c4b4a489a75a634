"""world_size-2 gloo test (CPU) of the host-side logic of the sharded path: table slices, the
all-to-all split arithmetic and the scatter / pack layouts used by dlrm_b200.dist.DistEngine
(reference: extend_distributed.py:47-62, :389-486)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dlrm_b200.dist import a2a_splits, pack_dT_into_send, scatter_recv_into_T, table_slices


def test_table_slices_match_reference_rule():
    # extend_distributed.get_my_slice: k, m = divmod(n, size); rank r gets k (+1 if r < m)
    assert table_slices(26, 8) == [(0, 4), (4, 8), (8, 11), (11, 14), (14, 17), (17, 20), (20, 23), (23, 26)]
    assert table_slices(5, 2) == [(0, 3), (3, 5)]
    assert table_slices(4, 4) == [(0, 1), (1, 2), (2, 3), (3, 4)]


def _f(b, k, d):  # value of pooled vector (global sample b, table k, column d)
    return 1000.0 * b + 10.0 * k + d


def _worker(rank, world, port, Tg, B, D, ok):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sl = table_slices(Tg, world)
        t0, t1 = sl[rank]
        Tl, Bg = t1 - t0, B * world
        send_splits, recv_splits = a2a_splits(Tg, world, rank, B, D)
        # forward: rank pools ITS tables for the GLOBAL batch
        send = torch.empty(Bg, Tl, D)
        for b in range(Bg):
            for t in range(Tl):
                for d in range(D):
                    send[b, t, d] = _f(b, t0 + t, d)
        recv = torch.empty(sum(recv_splits))
        dist.all_to_all_single(recv, send.view(-1), recv_splits, send_splits)
        Tbuf = torch.zeros(B, Tg + 1, D)
        scatter_recv_into_T(recv, Tbuf, Tg, world, B, D)
        for bl in range(B):
            for k in range(Tg):
                for d in range(D):
                    assert Tbuf[bl, 1 + k, d].item() == _f(rank * B + bl, k, d)
        assert Tbuf[:, 0, :].abs().sum().item() == 0  # dense feature slot untouched
        # backward: gradients of ALL tables for MY samples go back to the table owners
        dT = torch.zeros(B, Tg + 1, D)
        for bl in range(B):
            for k in range(Tg):
                for d in range(D):
                    dT[bl, 1 + k, d] = -_f(rank * B + bl, k, d)
        gsend = torch.empty(sum(recv_splits))
        pack_dT_into_send(dT, gsend, Tg, world, B, D)
        grecv = torch.empty(Bg * Tl * D)
        dist.all_to_all_single(grecv, gsend, send_splits, recv_splits)
        g = grecv.view(Bg, Tl, D)
        for b in range(Bg):
            for t in range(Tl):
                for d in range(D):
                    assert g[b, t, d].item() == -_f(b, t0 + t, d)
        # dense gradients: mean over ranks (DDP semantics)
        x = torch.full((7,), float(rank + 1))
        dist.all_reduce(x, op=dist.ReduceOp.SUM)
        x /= world
        assert torch.allclose(x, torch.full((7,), (world + 1) / 2.0))
        ok[rank] = 1
    finally:
        dist.destroy_process_group()


def test_a2a_layout_world2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    world = 2
    ctx = mp.get_context("spawn")
    ok = ctx.Array("i", [0] * world)
    mp.spawn(_worker, args=(world, port, 5, 4, 3, ok), nprocs=world, join=True)
    assert list(ok) == [1] * world


def test_push_route_equals_all_to_all_layout():
    """The peer-memory backward exchange (interact_bwd stores rows at push_route offsets) fills every
    owner's receive buffer exactly like pack_dT_into_send + all_to_all_single does (emulated on CPU)."""

    from dlrm_b200.dist import gather_route, push_route

    for world, n_tables, B, D in [(2, 5, 3, 4), (3, 7, 2, 2), (8, 26, 2, 4), (8, 11, 1, 2), (4, 2, 2, 2)]:
        sl = table_slices(n_tables, world)
        g = torch.Generator().manual_seed(world * 100 + n_tables)
        dT = [torch.randn(B, n_tables + 1, D, generator=g) for _ in range(world)]
        # all-to-all path: block d of rank s's send buffer lands as block s of rank d's receive buffer
        sends = []
        for s in range(world):
            buf = torch.zeros(B * n_tables * D)
            pack_dT_into_send(dT[s], buf, n_tables, world, B, D)
            sends.append(buf)
        want = []
        for d in range(world):
            t_d = sl[d][1] - sl[d][0]
            parts = []
            for s in range(world):
                o = sum(B * (e - b0) * D for b0, e in sl[:d])
                parts.append(sends[s][o:o + B * t_d * D])
            want.append(torch.cat(parts) if t_d else torch.zeros(0))
        # push path
        got = [torch.zeros(world * B * (e - s) * D) for s, e in sl]
        for s in range(world):
            route = push_route(n_tables, world, s, B, D)
            assert len(route) == n_tables
            for t, (owner, off, ld) in enumerate(route):
                assert sl[owner][0] <= t < sl[owner][1]
                for b in range(B):
                    got[owner][off + b * ld: off + b * ld + D] = dT[s][b, 1 + t, :]
        for d in range(world):
            assert torch.equal(got[d], want[d]), (world, n_tables, d)
            # the update kernel's view: sample g = s * B + b of owner d, local table k
            t_d = sl[d][1] - sl[d][0]
            if t_d:
                v = got[d].view(world * B, t_d, D)
                for s in range(world):
                    assert torch.equal(v[s * B:(s + 1) * B], dT[s][:, 1 + sl[d][0]:1 + sl[d][1], :])
        for r in range(world):
            assert gather_route(n_tables, world, r, D) == (1 + sl[r][0]) * D
