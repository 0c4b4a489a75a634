"""CPU port of the hot path on the reference's own dependency (torch ATen CPU kernels) -- TEST /
BASELINE INFRASTRUCTURE ONLY.  Used (a) as the timed `cpu_baseline` ("kind": "port") and the
`bench.py --impl reference` arm on the GPU box, where /root/reference does not exist, and (b) as a
second checker.  Never imported from dlrm_b200/.

The reference's arithmetic for this path IS torch: it calls
nn.EmbeddingBag(mode="sum", sparse=True) per table (dlrm_s_pytorch.py:277,452-457), nn.Linear /
ReLU / Sigmoid (:216-241), torch.cat / bmm / index / cat (:487-504), BCELoss/MSELoss (:385-393),
autograd (:1613) and optim/rwsadagrad.py or torch.optim.SGD (:1342-1369).  This file issues the
same ATen ops at the same granularity (one EmbeddingBag call per table, one addmm per layer, the
python-built li/lj index lists rebuilt every call, sparse COO embedding grads, coalesce() in the
optimizer), so its speed on the host cores is the speed of the reference's CPU path.
Pinned against the live-reference goldens by tests/test_oracle_golden.py::test_torch_port_*.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn


class CpuDLRM(nn.Module):
    def __init__(self, m_spa, ln_emb, ln_bot, ln_top, op="dot", itself=False, loss="bce",
                 loss_threshold=0.0, init="fast"):
        super().__init__()
        self.op, self.itself, self.thr = op, itself, loss_threshold
        self.emb_l = nn.ModuleList()
        first = {}
        for n in ln_emb:
            n = int(n)
            if init == "fast" and n in first:  # timing only needs realistic values, not distinct ones
                W = first[n].clone()
            else:
                a = float(np.sqrt(1.0 / n))
                W = torch.empty(n, m_spa).uniform_(-a, a)
                first[n] = W
            self.emb_l.append(nn.EmbeddingBag(n, m_spa, mode="sum", sparse=True, _weight=W))
        self.bot_l = self._mlp(ln_bot, -1)
        self.top_l = self._mlp(ln_top, len(ln_top) - 2)
        self.loss_fn = nn.BCELoss(reduction="mean") if loss == "bce" else nn.MSELoss(reduction="mean")

    @staticmethod
    def _mlp(ln, sigmoid_layer):
        layers = []
        for i in range(len(ln) - 1):
            n, m = int(ln[i]), int(ln[i + 1])
            L = nn.Linear(n, m, bias=True)
            with torch.no_grad():
                L.weight.normal_(0.0, float(np.sqrt(2.0 / (m + n))))
                L.bias.normal_(0.0, float(np.sqrt(1.0 / m)))
            layers += [L, nn.Sigmoid() if i == sigmoid_layer else nn.ReLU()]
        return nn.Sequential(*layers)

    def load(self, params):
        with torch.no_grad():
            for k, W in enumerate(params["emb"]):
                self.emb_l[k].weight.copy_(torch.as_tensor(W))
            for name, seq in (("bot", self.bot_l), ("top", self.top_l)):
                for i, (W, b) in enumerate(params[name]):
                    seq[2 * i].weight.copy_(torch.as_tensor(W))
                    seq[2 * i].bias.copy_(torch.as_tensor(b))

    def interact(self, x, ly):
        if self.op == "cat":
            return torch.cat([x] + ly, dim=1)
        B, d = x.shape
        T = torch.cat([x] + ly, dim=1).view((B, -1, d))
        Z = torch.bmm(T, torch.transpose(T, 1, 2))
        nf = T.shape[1]
        o = 1 if self.itself else 0
        li = torch.tensor([i for i in range(nf) for j in range(i + o)])  # rebuilt per call, as in
        lj = torch.tensor([j for i in range(nf) for j in range(i + o)])  # dlrm_s_pytorch.py:500-501
        return torch.cat([x, Z[:, li, lj]], dim=1)

    def forward(self, X, lS_o, lS_i):
        x = self.bot_l(X)
        ly = [E(lS_i[k], lS_o[k]) for k, E in enumerate(self.emb_l)]
        p = self.top_l(self.interact(x, ly))
        if 0.0 < self.thr < 1.0:
            p = torch.clamp(p, min=self.thr, max=1.0 - self.thr)
        return p


class RowWiseAdagradCPU:
    """optim/rwsadagrad.py:73-152 restated on torch sparse ops (CPU)."""

    def __init__(self, params, lr=0.01, eps=1e-10, lr_decay=0.0):
        self.params = list(params)
        self.lr, self.eps, self.lr_decay = lr, eps, lr_decay
        self.state = {}
        self.step_count = 0

    def zero_grad(self):
        for p in self.params:
            p.grad = None

    @torch.no_grad()
    def step(self):
        self.step_count += 1
        clr = self.lr / (1.0 + (self.step_count - 1.0) * self.lr_decay)
        for p in self.params:
            g = p.grad
            if g is None:
                continue
            if g.is_sparse:
                st = self.state.setdefault(p, torch.zeros(p.shape[0], dtype=torch.float32))
                g = g.coalesce()
                vals = g._values()
                if vals.numel() == 0:
                    continue
                # same ATen op sequence as the reference (sparse add_, sparse_mask, sparse add_)
                ind = g._indices()
                mom_upd = torch.sparse_coo_tensor(ind, vals.pow(2).mean(dim=1), (p.shape[0],))
                st.add_(mom_upd)
                stdv = st.sparse_mask(mom_upd.coalesce())._values().sqrt_().add_(self.eps)
                p.add_(torch.sparse_coo_tensor(ind, vals / stdv.view(-1, 1), p.shape), alpha=-clr)
            else:
                st = self.state.setdefault(p, torch.zeros_like(p))
                st.addcmul_(g, g, value=1.0)
                p.addcdiv_(g, st.sqrt().add_(self.eps), value=-clr)


def time_cpu_steps(model, opt, batches, nsteps, train=True):
    """Wall-clock seconds for nsteps fwd(+bwd+step) on pre-generated reference-format batches."""
    import time

    t0 = time.perf_counter()
    for s in range(nsteps):
        X, lS_o, lS_i, T = batches[s % len(batches)]
        if train:
            E = model.loss_fn(model(X, lS_o, lS_i), T)
            opt.zero_grad()
            E.backward()
            opt.step()
        else:
            with torch.no_grad():
                model(X, lS_o, lS_i)
    return time.perf_counter() - t0
