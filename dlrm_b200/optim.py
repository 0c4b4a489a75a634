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


class SGD(_Fused):
    _name = "sgd"


class RWSAdagrad(_Fused):
    _name = "rwsadagrad"
