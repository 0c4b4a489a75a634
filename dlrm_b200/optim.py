"""Fused optimizers for `dlrm_b200.DLRM_Net` with the constructor signatures the reference uses
(`opts[args.optimizer](parameters, lr=args.learning_rate)`, dlrm_s_pytorch.py:1342-1369):

    SGD(params, lr)          == torch.optim.SGD with sparse embedding gradients
    RWSAdagrad(params, lr)   == optim/rwsadagrad.py (row-wise sparse Adagrad; dense params: Adagrad)

`step()` launches the fused kernels (coalesce + row update in place, dense update + operand
refresh) on the gradients the last `backward()` left in the engine's buffers: no [nnz, D] sparse
gradient tensor exists.  The reference refuses Adagrad/RWSAdagrad on GPU (:1339-1340); here they run
on the device.  `param_groups[0]["lr"]` is honoured every step, so `LRPolicyScheduler` works.
"""
from __future__ import annotations

import torch



class _Fused(torch.optim.Optimizer):
    _name = "sgd"

    def __init__(self, params, lr=1e-2, lr_decay=0.0, weight_decay=0.0, initial_accumulator_value=0.0,
                 eps=1e-10):
        params = list(params)
        if weight_decay != 0.0:
            raise RuntimeError("weight_decay option is not compatible with sparse gradients")
        if initial_accumulator_value != 0.0:
            raise ValueError("initial_accumulator_value != 0 is not supported")
        super().__init__(params, dict(lr=lr, lr_decay=lr_decay, eps=eps))
        net = None
        for g in self.param_groups:
            for p in g["params"]:
                net = getattr(p, "_dlrm_net", None) or net
        if net is None:
            raise RuntimeError("dlrm_b200.optim optimizers take the parameters of a dlrm_b200.DLRM_Net")
        self.net = net() if callable(net) else net
        self.net._fused_opt = self
        self.net._engine.ensure_optimizer_state(self._name)

    def zero_grad(self, set_to_none: bool = True):
        self.net._pending = None      # discards a backward() whose step() was skipped
        for g in self.param_groups:
            for p in g["params"]:
                p.grad = None

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        net, eng = self.net, self.net._engine
        pend = net._pending
        if pend is None:
            return loss
        sp, linked = pend
        g = self.param_groups[0]
        eng.opt_step += 1
        clr = g["lr"] / (1.0 + (eng.opt_step - 1.0) * g["lr_decay"]) if self._name == "rwsadagrad" else g["lr"]
        eng.apply_optimizer(sp, self._name, clr, g["eps"], linked)
        net._pending = None
        return loss


    # ------------------------------------------------------------------ checkpoint (opt_state_dict)
    def _state_view(self, p):
        """('momentum', [rows] view) for an embedding table, ('sum', same-shape view) for an MLP parameter: the
        engine memory that holds the reference's per-parameter state (optim/rwsadagrad.py:86-100)."""
        eng = self.net._engine
        lo = eng.dense.data_ptr()
        if lo <= p.data_ptr() < lo + eng.dense.numel() * 4:
            off = (p.data_ptr() - lo) // 4
            return "sum", eng.dense_state[off:off + p.numel()].view(p.shape)
        for k in range(len(eng.row_base) - 1):
            if eng.table(k).data_ptr() == p.data_ptr():
                return "momentum", eng.momentum[int(eng.row_base[k]):int(eng.row_base[k + 1])]
        raise RuntimeError("parameter is not backed by the engine's memory")

    def state_dict(self):
        """torch.optim's layout with the reference optimizer's keys: RWSAdagrad keeps 'step' and, per parameter,
        'momentum' ([rows], embedding tables) or 'sum' (dense parameters); SGD has no state."""
        eng = self.net._engine
        params = [p for g in self.param_groups for p in g["params"]]
        state = {}
        if self._name == "rwsadagrad":
            for i, p in enumerate(params):
                kind, view = self._state_view(p)
                state[i] = {"step": int(eng.opt_step), kind: view.detach().clone()}
        group = {k: v for k, v in self.param_groups[0].items() if k != "params"}
        group["params"] = list(range(len(params)))
        return {"state": state, "param_groups": [group], "opt_step": int(eng.opt_step)}

    @torch.no_grad()
    def load_state_dict(self, sd):
        eng = self.net._engine
        params = [p for g in self.param_groups for p in g["params"]]
        for k, v in sd["param_groups"][0].items():
            if k != "params":
                self.param_groups[0][k] = v
        step = int(sd.get("opt_step", 0))
        for i, st in sd.get("state", {}).items():
            kind, view = self._state_view(params[int(i)])
            if kind not in st:
                raise KeyError("optimizer state of parameter %d has no '%s'" % (int(i), kind))
            view.copy_(torch.as_tensor(st[kind]).to(view.device))
            step = max(step, int(st.get("step", 0)))
        eng.opt_step = step


class SGD(_Fused):
    _name = "sgd"


class RWSAdagrad(_Fused):
    _name = "rwsadagrad"
