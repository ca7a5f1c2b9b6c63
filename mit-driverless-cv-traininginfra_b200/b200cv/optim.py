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
    """Host side of a fused step.  The per-step fast path is: list the parameters that have a gradient, form the key
    (param pointer, grad pointer) per tensor, look the chunk table up, bump ONE shared step counter, launch."""

    def _validate(self, ps):
        for p in ps:
            require_cuda(p, type(self).__name__)
            if p.dtype != torch.float32 or p.grad.dtype != torch.float32:
                raise TypeError(f"{type(self).__name__}: fp32 parameters and gradients only")
            if p.grad.is_sparse:
                raise RuntimeError(f"{type(self).__name__} does not support sparse gradients")
            if not p.is_contiguous() or not p.grad.is_contiguous():
                raise RuntimeError(f"{type(self).__name__}: parameters and gradients must be contiguous")

    def _ensure_state(self, ps, names):
        """State tensors are views of one flat arena per name; created on a parameter's first step.  All parameters
        that start together share ONE step-counter tensor (torch keeps one per parameter; sharing makes the per-step
        host cost independent of the number of tensors and reads the same through state_dict())."""
        new = [p for p in ps if "step" not in self.state[p]]
        if not new:
            return
        total = sum((p.numel() + 3) // 4 * 4 for p in new)
        for n in names:
            arena = torch.zeros(total, dtype=torch.float32, device=new[0].device)
            off = 0
            for p in new:
                self.state[p][n] = arena[off:off + p.numel()].view_as(p)
                off += (p.numel() + 3) // 4 * 4
        step = torch.tensor(0.0)
        for p in new:
            self.state[p]["step"] = step

    def _bump_steps(self, ps):
        """Increment every distinct step tensor once; returns {step value: [params]}."""
        seen, by_step = {}, {}
        for p in ps:
            t = self.state[p]["step"]
            if id(t) not in seen:
                t += 1
                seen[id(t)] = int(t)
            by_step.setdefault(seen[id(t)], []).append(p)
        return by_step

    def _table(self, ps, names):
        """Device chunk table for these parameters, cached until a parameter or gradient pointer changes."""
        key = tuple([(p.data_ptr(), p.grad.data_ptr()) for p in ps])
        cache = self.__dict__.setdefault("_b200cv_tables", {})
        table = cache.get(key)
        if table is None:
            self._validate(ps)
            rows = []
            for p in ps:
                n, pp, gp = p.numel(), p.data_ptr(), p.grad.data_ptr()
                sp = [self.state[p][nm].data_ptr() for nm in names] + [0, 0]
                for o in range(0, n, CHUNK):
                    rows.append((pp + 4 * o, gp + 4 * o, sp[0] + 4 * o if sp[0] else 0, sp[1] + 4 * o if sp[1] else 0,
                                 min(CHUNK, n - o)))
            table = torch.tensor(rows, dtype=torch.int64).reshape(-1, 5).to(ps[0].device)
            if len(cache) >= 8:
                cache.clear()
            cache[key] = table
        return table

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self.__dict__.pop("_b200cv_tables", None)  # the loaded state lives in new tensors


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
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            self._ensure_state(ps, names)
            b1, b2 = group["betas"]
            # parameters that joined later carry their own step count: one launch per distinct count
            for t, plist in self._bump_steps(ps).items():
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
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            names = ("momentum_buffer",) if group["momentum"] != 0 else ()
            self._ensure_state(ps, names)  # zero-filled buffers: mu*0 + g = g is torch's first step exactly
            self._bump_steps(ps)
            table = self._table(ps, names)
            lib().call("b200cv_sgd_step_multi", ptr(table), table.shape[0], float(group["lr"]),
                       float(group["momentum"]), float(group["weight_decay"]), 0, stream_ptr())
        return loss
