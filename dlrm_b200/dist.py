"""One-process-per-GPU DLRM: table-wise sharded embeddings + data-parallel MLPs.

Replaces `DLRM_Net.distributed_forward` (dlrm_s_pytorch.py:528-585) and `extend_distributed.py`
(`get_my_slice` :47-51, `get_split_lengths` :54-62, `alltoall` :541-576 and its autograd pair
`All2All_Req/Wait` :389-486, DDP of the MLPs :1329-1336):

  * tables are split into contiguous slices over the ranks exactly like `get_my_slice`;
  * every rank pools ITS tables for the GLOBAL batch (one gather launch), the pooled vectors are
    exchanged so that every rank ends up with ALL tables for ITS batch slice, inside the
    interaction operand T (forward), and the per-bag gradients travel the opposite way (backward)
    into the fused coalesce + row-wise-Adagrad update of the owning rank;
  * the MLPs are replicated; their gradients are averaged with one NCCL all-reduce (DDP semantics:
    mean over ranks of the local-mean-loss gradients; embedding gradients are NOT averaged -- the
    reference's all-to-all backward simply routes them, `extend_distributed.py:467-486`).

Exchange back ends
  "nccl" : `all_to_all_single` on packed send/recv buffers (the reference's collective).
  "p2p"  : the gather kernel stores each pooled row directly into the owner rank's T buffer
           through peer-mapped memory (NVLink stores) and the update kernel reads the dY rows
           from the peers' dT buffers -- no staging buffers, no separate collective.

Launch: `torchrun --nproc-per-node N ...` (env RANK/LOCAL_RANK/WORLD_SIZE/MASTER_*).
"""
from __future__ import annotations

import json
import os
import time
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


# ---------------------------------------------------------------------------- host-side layout logic
def table_slices(n_tables: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous [start, end) table slice of every rank (extend_distributed.get_my_slice)."""
    k, m = divmod(n_tables, world)
    return [(r * k + min(r, m), (r + 1) * k + min(r + 1, m)) for r in range(world)]


def a2a_splits(n_tables: int, world: int, rank: int, local_batch: int, dim: int):
    """(send_splits, recv_splits) in ELEMENTS for the forward exchange of pooled vectors.

    send buffer of rank r : [world][local_batch][T_r][dim]   (sample-major: block d goes to rank d)
    recv buffer of rank r : [world][local_batch][T_s][dim]   (block s came from rank s)
    The backward exchange uses the same numbers with the roles swapped."""
    sl = table_slices(n_tables, world)
    t_mine = sl[rank][1] - sl[rank][0]
    send = [local_batch * t_mine * dim] * world
    recv = [local_batch * (e - s) * dim for s, e in sl]
    return send, recv


def scatter_recv_into_T(recv: torch.Tensor, Tbuf: torch.Tensor, n_tables: int, world: int,
                        local_batch: int, dim: int):
    """recv [sum_s B*T_s*D] -> Tbuf[:B, 1 + start_s : 1 + end_s, :] for every source rank s."""
    o = 0
    for s, e in table_slices(n_tables, world):
        n = local_batch * (e - s) * dim
        if e > s:
            Tbuf[:local_batch, 1 + s:1 + e, :].copy_(recv[o:o + n].view(local_batch, e - s, dim))
        o += n


def pack_dT_into_send(dT: torch.Tensor, gsend: torch.Tensor, n_tables: int, world: int,
                      local_batch: int, dim: int):
    """dT[:B, 1 + start_d : 1 + end_d, :] -> gsend block d (gradients of rank d's tables)."""
    o = 0
    for s, e in table_slices(n_tables, world):
        n = local_batch * (e - s) * dim
        if e > s:
            gsend[o:o + n].view(local_batch, e - s, dim).copy_(dT[:local_batch, 1 + s:1 + e, :])
        o += n


def push_route(n_tables: int, world: int, rank: int, local_batch: int, dim: int):
    """Where rank `rank`'s interact_bwd stores the gradient rows of every table (peer-memory exchange).

    Returns one (owner, offset, ld) per table t, in ELEMENTS: the row of local sample b goes to
    `recv_of(owner)[offset + b * ld : ... + dim]`.  The receive buffer of a rank with T_r tables is
    [world][local_batch][T_r][dim]; slab `rank` holds what this rank sends, so the owner's update kernel
    sees the same [global batch][T_r][dim] view the all-to-all path assembles (a2a_splits)."""
    out = []
    for r, (s, e) in enumerate(table_slices(n_tables, world)):
        for t in range(s, e):
            out.append((r, (rank * local_batch * (e - s) + (t - s)) * dim, (e - s) * dim))
    return out


def gather_route(n_tables: int, world: int, rank: int, dim: int):
    """Forward twin of push_route: byte-free description of where the gather of rank `rank` stores the
    pooled row of (global sample g, local table k): T_of(g // local_batch)[(g % local_batch), 1 + s + k, :]
    -> returns the element offset of feature (1 + s) inside one sample of T ([F, dim] per sample)."""
    s, _ = table_slices(n_tables, world)[rank]
    return (1 + s) * dim


def init_distributed(backend: Optional[str] = None):
    """Process-group bring-up from the torchrun environment (extend_distributed.init_distributed)."""
    if dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local)
        dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
    else:
        dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world


# ---------------------------------------------------------------------------- the sharded engine
class DistEngine:
    def __init__(self, m_spa: int, ln_emb: Sequence[int], ln_bot: Sequence[int], ln_top: Sequence[int], *,
                 local_batch: int, device=None, gemm: str = "tc", loss: str = "bce", exchange: str = "nccl",
                 **kw):
        from .engine import Engine

        if gemm == "simt":
            raise SystemExit("ERROR: dlrm_b200.dist runs on the tensor-core path (gemm='tc' or 'tc_bf16')")
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.Tg = len(ln_emb)
        if self.Tg < self.world:
            raise SystemExit("ERROR: only (%d) sparse features for (%d) devices, table partitions will fail"
                             % (self.Tg, self.world))
        self.D = int(m_spa)
        self.B = int(local_batch)
        self.Bg = self.B * self.world
        self.slices = table_slices(self.Tg, self.world)
        self.t0, self.t1 = self.slices[self.rank]
        self.Tl = self.t1 - self.t0
        self.exchange = exchange
        if device is None:
            device = "cuda:%d" % torch.cuda.current_device()
        self.device = torch.device(device)
        self.eng = Engine(m_spa, list(ln_emb[self.t0:self.t1]), ln_bot, ln_top, n_features=self.Tg + 1,
                          loss=loss, device=device, max_batch=self.B, gemm=gemm,
                          sigmoid_top=len(ln_top) - 2, **kw)
        e = self.eng
        f32 = torch.float32
        self.send_splits, self.recv_splits = a2a_splits(self.Tg, self.world, self.rank, self.B, self.D)
        self.send = torch.zeros(self.Bg * self.Tl * self.D, dtype=f32, device=self.device)
        self.recv = torch.zeros(sum(self.recv_splits), dtype=f32, device=self.device)
        self.gsend = torch.zeros(sum(self.recv_splits), dtype=f32, device=self.device)
        self.grecv = torch.zeros(self.Bg * self.Tl * self.D, dtype=f32, device=self.device)
        e.gather_fn, e.update_fn, e.dense_sync_fn = self._gather, self._update, self._dense_sync
        self._flag = torch.zeros(1, dtype=f32, device=self.device)
        if exchange == "p2p":
            self._setup_p2p()
            e.gather_fn, e.update_fn = self._gather_p2p, self._update_p2p
        elif exchange != "nccl":
            raise SystemExit("ERROR: exchange must be nccl or p2p")

    # ------------------------------------------------------------------ peer-mapped exchange
    def _setup_p2p(self):
        """Map every rank's interaction operand T and gradient receive buffer into this process (CUDA IPC
        over NVLink).  Both directions PUSH: the gather stores pooled rows into the T of the rank that owns
        the sample, interact_bwd stores per-table gradient rows into the receive buffer of the table's owner;
        every load on the data path stays local (L2-cacheable)."""
        import ctypes as C

        e = self.eng
        e._tc_prepare(self.B)    # final allocation of the dense-gradient arena (split-K slabs)
        # two barrier channels (slots + epoch each): 0 = embedding stream (gather / update), 1 = main
        # stream (dense-gradient sync).  Barriers of one channel are ordered by their stream on every
        # rank; the two streams may interleave differently across ranks, so they must not share slots.
        self._sig = torch.zeros(32, dtype=torch.int32, device=self.device)
        self._epoch = torch.zeros(2, dtype=torch.int32, device=self.device)
        torch.cuda.synchronize()
        from . import _lib as _l

        def export(t):
            # handle of the cudaMalloc block holding t + t's byte offset inside it
            h = C.create_string_buffer(64)
            off = C.c_int64()
            _l.check(e.lib.dlrm_b200_ipc_export(t.data_ptr(), h, C.byref(off)), "ipc_export")
            return h.raw, int(off.value)

        # gradient receive buffer: slab s = [B, Tl, D] written by rank s's interact_bwd (push over NVLink)
        self._grecv_p2p = torch.zeros(self.world * self.B * max(self.Tl, 1) * self.D, dtype=torch.float32,
                                      device=self.device)
        torch.cuda.synchronize()
        mine = tuple(export(t) for t in (e.Tbuf, self._grecv_p2p, self._sig, e.dense_grad))
        allh = [None] * self.world
        dist.all_gather_object(allh, mine)
        self._ipc_bases = {}
        pT, pdT, psig, pgrad = [], [], [], []
        col = gather_route(self.Tg, self.world, self.rank, self.D) * 4

        def imp(handle, offset):
            if handle not in self._ipc_bases:       # one open per exported allocation
                base = C.c_void_p()
                _l.check(e.lib.dlrm_b200_ipc_open(handle, self.device.index, C.byref(base)), "ipc_open")
                self._ipc_bases[handle] = base.value
            return self._ipc_bases[handle] + offset

        for r in range(self.world):
            if r == self.rank:
                ptrs = [e.Tbuf.data_ptr(), self._grecv_p2p.data_ptr(), self._sig.data_ptr(), e.dense_grad.data_ptr()]
            else:
                ptrs = [imp(hd, off) for hd, off in allh[r]]
            pT.append(ptrs[0] + col)
            pdT.append(ptrs[1])
            psig.append(ptrs[2])
            pgrad.append(ptrs[3])
        W = self.world
        self._peer_T = (C.c_void_p * W)(*pT)
        # update side: slab s of MY receive buffer holds the gradients of rank s's samples
        slab = self.B * self.Tl * self.D * 4
        self._peer_dT = (C.c_void_p * W)(*[self._grecv_p2p.data_ptr() + s_ * slab for s_ in range(W)])
        # interact_bwd side: feature 1 + t -> slab `rank` of the owner of table t; feature 0 stays local
        dst, ld = [e.dT.data_ptr()], [e.F * self.D]
        for owner, off, ldr in push_route(self.Tg, W, self.rank, self.B, self.D):
            dst.append(pdT[owner] + off * 4)
            ld.append(ldr)
        assert len(dst) == e.F
        e.dT_route = ((C.c_void_p * e.F)(*dst), (C.c_int64 * e.F)(*ld))
        self._peer_sig = [(C.c_void_p * W)(*[p + 64 * ch for p in psig]) for ch in range(2)]
        self._peer_grad = (C.c_void_p * W)(*pgrad)
        self.own_sync = os.environ.get("DLRM_P2P_NCCL_SYNC") != "1"   # our kernels instead of NCCL all_reduce
        if self.own_sync:
            e.dense_sync_fn = self._dense_sync_p2p
        dist.barrier()

    def _barrier(self, channel: int = 0):
        """Device-side ordering across ranks on the current stream (no host sync)."""
        if getattr(self, "own_sync", False):
            from . import _lib
            from .engine import _stream

            _lib.check(self.eng.lib.dlrm_b200_p2p_barrier(self._peer_sig[channel], self.rank, self.world,
                                                          self._epoch.data_ptr() + 4 * channel, _stream()),
                       "p2p_barrier")
            self.eng.n_launch += 1
        else:
            dist.all_reduce(self._flag)

    def _dense_sync_p2p(self):
        """Mean of the dense gradients over the ranks with our own two-shot all-reduce over NVLink."""
        from . import _lib
        from .engine import _stream

        self._barrier(1)      # every rank's gradient arena is complete
        _lib.check(self.eng.lib.dlrm_b200_p2p_allreduce_mean(self._peer_grad, self.rank, self.world,
                                                             self.eng.dense_numel, _stream()), "p2p_allreduce_mean")
        self.eng.n_launch += 1
        self._barrier(1)      # every slice has been written back everywhere

    def _gather_p2p(self, sp, link):
        from . import _lib
        from .engine import _stream

        e = self.eng
        if not link:
            self._barrier()   # inference loops: peers must be done reading T of the previous batch
        FD = e.F * self.D
        if link:
            e._ensure_link(sp.nnz_total if sp.include_last else sum(int(i.numel()) for i in sp.indices))
        desc = e._fwd_desc(sp, range(self.Tl))
        bdesc = e._bwd_desc_chunk(sp, list(range(self.Tl)))[0] if link else None
        import ctypes as C

        filt = link and e.use_filter
        if filt:
            e.filter.zero_()
        _lib.check(e.lib.dlrm_b200_emb_bag_fwd_p2p(desc, bdesc, self.Tl, self.D, sp.batch, sp.idx_bytes,
                                                   int(sp.include_last), e.link.data_ptr() if link else None,
                                                   self._peer_T, self.world, self.B, FD, self.D,
                                                   C.byref(e.dedup) if filt else None, _stream()),
                   "emb_bag_fwd_p2p")
        e.n_launch += 1
        if link:
            e._filtered = filt
            if filt:
                e.emb_classify(sp)
        self._barrier()       # every rank's pooled rows have landed in every T

    def _update_p2p(self, sp, optimizer, clr):
        from . import _lib
        from .engine import _OPT, _stream

        import ctypes as C

        e = self.eng
        self._barrier()       # every rank's interact_bwd stores have landed in my receive buffer
        bdesc, _ = e._bwd_desc_chunk(sp, list(range(self.Tl)))
        _lib.check(e.lib.dlrm_b200_emb_bwd_update_p2p(bdesc, self.Tl, self.D, sp.batch, sp.idx_bytes,
                                                      int(sp.include_last), e.link.data_ptr(), self._peer_dT,
                                                      self.world, self.B, self.Tl * self.D, self.D,
                                                      _OPT[optimizer], clr,
                                                      1e-10, C.byref(e.dedup) if e._filtered else None, _stream()),
                   "emb_bwd_update_p2p")
        e.n_launch += 1

    # -- forward: pool local tables for the global batch, exchange, land in T
    def _gather(self, sp, link):
        e = self.eng
        e.emb_forward(sp, self.send, self.Tl * self.D, self.D, link)
        dist.all_to_all_single(self.recv, self.send, self.recv_splits, self.send_splits)
        scatter_recv_into_T(self.recv, e.Tbuf, self.Tg, self.world, self.B, self.D)

    # -- backward: route per-bag gradients to the table owners, fused coalesce + optimizer there
    def _update(self, sp, optimizer, clr):
        e = self.eng
        pack_dT_into_send(e.dT, self.gsend, self.Tg, self.world, self.B, self.D)
        dist.all_to_all_single(self.grecv, self.gsend, self.send_splits, self.recv_splits)
        e.emb_update(sp, self.grecv, self.Tl * self.D, self.D, optimizer, clr)

    def _dense_sync(self):
        dist.all_reduce(self.eng.dense_grad[:self.eng.dense_numel], op=dist.ReduceOp.AVG)

    def sync_dense_params_from_rank0(self):
        """DDP broadcasts rank 0's MLP weights at wrap time (SURVEY A: reference quirk)."""
        dist.broadcast(self.eng.dense, src=0)
        self.eng.mark_params_changed()

    def forward(self, X_local, sp_global_local_tables):
        return self.eng.forward(X_local, sp_global_local_tables)

    def train_step(self, X_local, sp_global_local_tables, target_local, lr, optimizer="rwsadagrad"):
        return self.eng.train_step(X_local, sp_global_local_tables, target_local, lr, optimizer)


# ---------------------------------------------------------------------------- synthetic sharded batches
def make_sharded_batch(step_seed: int, ln_emb: Sequence[int], rank: int, world: int, local_batch: int,
                       m_den: int = 13, lmax: int = 10, pin: bool = True):
    """Rank-local view of one GLOBAL synthetic batch: indices of the rank's tables for all
    world*local_batch samples (packed format) + the rank's slice of dense features and targets.
    Every table / slice has its own seed, so all ranks agree on the global batch."""
    from .data import HostBatch, PackedLayout, fill_batch

    t0, t1 = table_slices(len(ln_emb), world)[rank]
    rows = list(ln_emb[t0:t1])
    Bg = local_batch * world
    cap = int(Bg * sum(min(int(r), lmax) for r in rows))
    hb = HostBatch(PackedLayout(Bg, len(rows), m_den, cap), pin)
    fill_batch(hb, np.random.default_rng([step_seed, 7, rank]), rows, lmax)
    rng = np.random.default_rng([step_seed, 11])
    Xg = rng.random((Bg, m_den), dtype=np.float32)
    Tg = np.round(rng.random((Bg, 1), dtype=np.float32))
    sl = slice(rank * local_batch, (rank + 1) * local_batch)
    X = torch.from_numpy(Xg[sl].copy())
    T = torch.from_numpy(Tg[sl].copy())
    if pin and torch.cuda.is_available():
        X, T = X.pin_memory(), T.pin_memory()
    return hb, X, T


# ---------------------------------------------------------------------------- bench entry (bench.py --gpus N)
def bench_main(args, CFG, metric_name, config_dict, ClockSampler):
    from .data import DeviceBatch

    rank, world = init_distributed("nccl")
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    dev = "cuda:%d" % local
    torch.cuda.set_device(local)
    train = args.workload != "cfg1"
    D, T, B = CFG["m_spa"], CFG["T"], CFG["B"]
    ln_emb = [CFG["rows"]] * T
    ln_top = [D + (T + 1) * T // 2] + CFG["top_tail"]
    exchange = getattr(args, "exchange", "auto")
    if exchange == "auto":
        # capability check, same answer on every rank (one node): peer access between all GPU pairs
        n = torch.cuda.device_count()
        ok = world <= n and all(torch.cuda.can_device_access_peer(a, b)
                                for a in range(world) for b in range(world) if a != b)
        exchange = "p2p" if ok else "nccl"
    de = DistEngine(D, ln_emb, CFG["ln_bot"], ln_top, local_batch=B, device=dev, gemm=args.gemm,
                    exchange=exchange)
    de.eng.init_params(100 + rank)
    de.sync_dense_params_from_rank0()
    de.eng.ensure_optimizer_state("rwsadagrad")
    ring = []
    for i in range(args.ring):
        hb, X, Tt = make_sharded_batch(1000 + i, ln_emb, rank, world, B, 13, CFG["lmax"])
        db = DeviceBatch(hb.layout, dev)
        db.load(hb, non_blocking=False)
        ring.append((hb, db, X, X.to(dev), Tt, Tt.to(dev)))
    lr = 0.01

    # K static staging buffers per rank; with the NCCL-free exchange the whole sharded step (K of them per
    # graph, update of step j overlapping step j+1) is captured in one CUDA graph per rank
    import types

    Kp = 1
    want_graph = (de.exchange == "p2p" and getattr(de, "own_sync", False)) or os.environ.get("DLRM_DIST_GRAPH") == "1"
    want_graph = want_graph and not getattr(args, "no_graph", False)
    if train and not getattr(args, "no_pipeline", False):
        cand = getattr(args, "pipeline", 1)
        if cand >= 1 and args.steps % cand == 0:
            Kp = cand
    stages = []
    for j in range(Kp):
        st = DeviceBatch(ring[0][0].layout, dev)
        st.load(ring[0][0], non_blocking=False)
        stages.append(types.SimpleNamespace(db=st, sparse=st.sparse, X=ring[0][3].clone(), target=ring[0][5].clone()))
    graph = None
    if want_graph:
        from .engine import GraphedTrainStep, GraphedTrainSteps

        try:
            if train:
                graph = GraphedTrainSteps(de.eng, stages, lr, "rwsadagrad")
            else:
                graph = GraphedTrainStep(de.eng, stages[0], lr, "rwsadagrad", train=False)
        except Exception as ex:  # noqa: BLE001
            if rank == 0:
                print("dist: CUDA-graph capture failed (%s); running eagerly" % str(ex)[:200], flush=True)
            graph = None

    def run_round():
        if graph is not None:
            return graph.replay()
        out = None
        for j, st in enumerate(stages):
            if train:
                out = de.eng.train_step(st.X, st.sparse, st.target, lr, "rwsadagrad", join_update=(j == Kp - 1))
            else:
                out = de.eng.forward(st.X, st.sparse)
        return out

    def resident_round(i):
        for j, st in enumerate(stages):
            hb, db, Xh, Xd, Th, Td = ring[(i * Kp + j) % args.ring]
            nbytes = db.layout.used(db.nnz)
            st.db.buf[:nbytes].copy_(db.buf[:nbytes], non_blocking=True)
            st.X.copy_(Xd, non_blocking=True)
            st.target.copy_(Td, non_blocking=True)
        return run_round()

    rounds, wrounds = args.steps // Kp, max((args.warmup + Kp - 1) // Kp, 1)
    for w in range(wrounds):
        resident_round(w)
    torch.cuda.synchronize()
    dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    n0 = de.eng.n_launch
    ev0.record()
    for r in range(rounds):
        resident_round(wrounds + r)
    ev1.record()
    torch.cuda.synchronize()
    dist.barrier()
    t1 = time.time()
    launches = de.eng.n_launch - n0
    ms = torch.tensor([ev0.elapsed_time(ev1) / args.steps], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())

    # e2e: host batches (packed sparse part + dense slice) copied every step, loss read back
    loss_host = torch.zeros(1).pin_memory()
    h2d = 0

    def e2e_round(i):
        nonlocal h2d
        for j, st in enumerate(stages):
            hb, db, Xh, Xd, Th, Td = ring[(i * Kp + j) % args.ring]
            h2d += st.db.load(hb)
            st.X.copy_(Xh, non_blocking=True)
            st.target.copy_(Th, non_blocking=True)
            h2d += Xh.numel() * 4 + Th.numel() * 4
        out = run_round()
        loss_host.copy_(out.view(-1)[-1:], non_blocking=True)

    for w in range(2):
        e2e_round(w)
    h2d = 0
    torch.cuda.synchronize()
    dist.barrier()
    ev0.record()
    for r in range(rounds):
        e2e_round(r)
    ev1.record()
    torch.cuda.synchronize()
    dist.barrier()
    ms2 = torch.tensor([ev0.elapsed_time(ev1) / args.steps], device=dev)
    dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    ms2 = float(ms2.item())
    if rank == 0:
        clocks = sampler.stop(t0, t1)
        Bg = B * world
        line = {
            "metric": metric_name(train), "value": Bg / (ms * 1e-3), "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": {"simt": "fp32", "tc": "fp32 (bf16x3 split on tcgen05, fp32 accumulate)",
                      "tc_bf16": "bf16"}[args.gemm],
            "data": "synthetic", "config": config_dict(args, world),
            "roofline": None, "cpu_baseline": None,
            "e2e": {"value": Bg / (ms2 * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": int(h2d / args.steps),
                    "d2h_bytes_per_step": 4, "ms_per_step": ms2,
                    "note": "per rank: packed pinned sparse batch (its tables, global batch) + dense slice, "
                            "H2D every step, loss read back"},
            "gpu_launches": int(launches), "exchange": de.exchange, "cuda_graph": graph is not None, "steps_per_graph": Kp,
            "a2a_bytes_per_rank_per_step": int(2 * 4 * (sum(de.send_splits) - de.send_splits[rank])),
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    dist.barrier()
    dist.destroy_process_group()
