"""Fused optimizer steps: drop-in ``torch.optim.Optimizer`` subclasses for the two optimizers the
reference's training scripts build right after backward --

    torch.optim.Adam(params, lr, weight_decay)            CVC-YOLOv3/train.py:181, RektNet/train_eval.py:263
    torch.optim.SGD(params, lr, momentum, weight_decay)   CVC-YOLOv3/train.py:185

-- with one kernel launch per param group (b200cv_adam_step_multi / b200cv_sgd_step_multi) instead of one chain of
element-wise kernels per parameter.  ``param_groups`` / ``state`` keep torch's layout (``step``, ``exp_avg``,
``exp_avg_sq`` / ``momentum_buffer``), so LR schedulers (StepLR, ExponentialLR: train.py:199, train_eval.py:264) and
``state_dict()`` work unchanged.  Parameters and gradients stay wherever they live (e.g. views of the engine's flat
gradient arena); optimizer state lives in flat arenas carved per parameter.
"""
from __future__ import annotations

import math

import torch

from .lib import lib, ptr, require_cuda, stream_ptr

CHUNK = 16384  # elements per CTA (a multiple of 4: chunk starts keep the 16-byte alignment of the tensors)


class _FusedBase(torch.optim.Optimizer):
    _n_state = 0

    def _group_tensors(self, group):
        ps = [p for p in group["params"] if p.grad is not None]
        for p in ps:
            require_cuda(p, type(self).__name__)
            if p.dtype != torch.float32 or p.grad.dtype != torch.float32:
                raise TypeError(f"{type(self).__name__}: fp32 parameters and gradients only")
            if p.grad.is_sparse:
                raise RuntimeError(f"{type(self).__name__} does not support sparse gradients")
            if not p.is_contiguous():
                raise RuntimeError(f"{type(self).__name__}: parameters must be contiguous")
        return ps

    def _ensure_state(self, group, ps, names):
        """State tensors are views of one flat arena per (group, name); created on a parameter's first step."""
        new = [p for p in ps if not all(n in self.state[p] for n in names)]
        if not new:
            return
        total = sum((p.numel() + 3) // 4 * 4 for p in new)
        for n in names:
            arena = torch.zeros(total, dtype=torch.float32, device=new[0].device)
            off = 0
            for p in new:
                self.state[p][n] = arena[off:off + p.numel()].view_as(p)
                off += (p.numel() + 3) // 4 * 4
        for p in new:
            self.state[p].setdefault("step", torch.tensor(0.0))

    def _table(self, ps, names):
        """Device chunk table, cached until a pointer changes (e.g. a new gradient arena)."""
        key = []
        for p in ps:
            if not p.grad.is_contiguous():
                p.grad = p.grad.contiguous()
            key.append((p.data_ptr(), p.grad.data_ptr(), p.numel()) + tuple(self.state[p][n].data_ptr() for n in names))
        key = tuple(key)
        cache = self.__dict__.setdefault("_b200cv_tables", {})
        table = cache.get(key)
        if table is None:
            rows = []
            for pp, gp, n, *sp in key:
                sp = list(sp) + [0] * (2 - len(sp))
                for o in range(0, n, CHUNK):
                    rows.append((pp + 4 * o, gp + 4 * o, sp[0] + 4 * o if sp[0] else 0, sp[1] + 4 * o if sp[1] else 0,
                                 min(CHUNK, n - o)))
            table = torch.tensor(rows, dtype=torch.int64).reshape(-1, 5).to(ps[0].device)
            if len(cache) >= 8:
                cache.clear()
            cache[key] = table
        return table


class FusedAdam(_FusedBase):
    """torch.optim.Adam(params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0) -- same update rule."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("FusedAdam: invalid hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        names = ("exp_avg", "exp_avg_sq")
        for group in self.param_groups:
            ps = self._group_tensors(group)
            if not ps:
                continue
            self._ensure_state(group, ps, names)
            b1, b2 = group["betas"]
            # parameters that joined later have their own step count: one launch per distinct count
            by_step = {}
            for p in ps:
                self.state[p]["step"] += 1
                by_step.setdefault(int(self.state[p]["step"]), []).append(p)
            for t, plist in by_step.items():
                table = self._table(plist, names)
                bc1 = 1.0 - b1 ** t
                bc2 = 1.0 - b2 ** t
                lib().call("b200cv_adam_step_multi", ptr(table), table.shape[0], float(group["lr"] / bc1), float(b1),
                           float(b2), float(group["eps"]), float(group["weight_decay"]), float(math.sqrt(bc2)),
                           stream_ptr())
        return loss


class FusedSGD(_FusedBase):
    """torch.optim.SGD(params, lr, momentum=0, weight_decay=0) (dampening 0, no Nesterov) -- same update rule."""

    def __init__(self, params, lr=1e-3, momentum=0.0, weight_decay=0.0):
        if lr < 0 or momentum < 0 or weight_decay < 0:
            raise ValueError("FusedSGD: invalid hyper-parameter")
        super().__init__(params, dict(lr=lr, momentum=momentum, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            ps = self._group_tensors(group)
            if not ps:
                continue
            names = ("momentum_buffer",) if group["momentum"] != 0 else ()
            self._ensure_state(group, ps, names)  # zero-filled buffers: mu*0 + g = g is torch's first step exactly
            for p in ps:
                self.state[p]["step"] = self.state[p].get("step", torch.tensor(0.0)) + 1
            table = self._table(ps, names)
            lib().call("b200cv_sgd_step_multi", ptr(table), table.shape[0], float(group["lr"]),
                       float(group["momentum"]), float(group["weight_decay"]), 0, stream_ptr())
        return loss
