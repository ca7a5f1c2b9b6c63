"""ORACLE (test infrastructure, never shipped, never on the product path).

CPU fp32 restatement of the RektNet hot path (KeypointNet + CrossRatioLoss) of
cv-core/MIT-Driverless-CV-TrainingInfra as pure functions over a dict of named tensors.  Parity is
PINNED against vectors made by importing the reference (oracle/gen_golden.py -> tests/golden/).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

NET_SIZE = 16  # RektNet/keypoint_net.py:15
BLOCKS = [("res1", 16, 16), ("res2", 16, 32), ("res3", 32, 64), ("res4", 64, 128)]  # :21-24


def init_params(num_kpt: int = 7, seed: int = 0) -> Tuple[Dict[str, torch.Tensor], Dict[str, torch.Tensor]]:
    """State-dict names of KeypointNet (keypoint_net.py:17-25, resnet.py:12-20) with the
    _initialize_weights recipe (:33-44): Kaiming-normal fan_out convs, zero biases, BN gamma=1 beta=0."""
    g = torch.Generator().manual_seed(seed)
    params: Dict[str, torch.Tensor] = {}
    buffers: Dict[str, torch.Tensor] = {}

    def conv(name, cin, cout, k):
        std = math.sqrt(2.0 / (cout * k * k))
        params[name + ".weight"] = torch.empty(cout, cin, k, k).normal_(0.0, std, generator=g)
        params[name + ".bias"] = torch.zeros(cout)

    def bn(name, c):
        params[name + ".weight"] = torch.ones(c)
        params[name + ".bias"] = torch.zeros(c)
        buffers[name + ".running_mean"] = torch.zeros(c)
        buffers[name + ".running_var"] = torch.ones(c)

    conv("conv", 3, NET_SIZE, 7)
    bn("bn", NET_SIZE)
    for name, cin, cout in BLOCKS:
        conv(f"{name}.conv1", cin, cout, 3)
        bn(f"{name}.bn1", cout)
        conv(f"{name}.conv2", cout, cout, 3)
        bn(f"{name}.bn2", cout)
        conv(f"{name}.shortcut_conv", cin, cout, 1)
        bn(f"{name}.shortcut_bn", cout)
    conv("out", NET_SIZE * 8, num_kpt, 1)
    return params, buffers


class _StoreBf16(torch.autograd.Function):
    """Straight-through bf16 storage rounding (only for the `emulate_bf16` diagnostic mode)."""

    @staticmethod
    def forward(ctx, t):
        return t.to(torch.bfloat16).float()

    @staticmethod
    def backward(ctx, g):
        return g


_ident = lambda t: t


def _bn(x, params, buffers, name, training):
    return F.batch_norm(x, buffers[name + ".running_mean"], buffers[name + ".running_var"], params[name + ".weight"],
                        params[name + ".bias"], training=training, momentum=0.1, eps=1e-5)


def res_block(x, params, buffers, name, training, q=_ident):
    """RektNet/resnet.py:22-27: relu(BN(1x1(x)) + BN(3x3(relu(BN(3x3 dil2 pad2 (x))))))."""
    c1 = q(F.conv2d(x, q(params[f"{name}.conv1.weight"]), params[f"{name}.conv1.bias"], padding=2, dilation=2))
    a1 = q(F.relu(_bn(c1, params, buffers, f"{name}.bn1", training)))
    c2 = q(F.conv2d(a1, q(params[f"{name}.conv2.weight"]), params[f"{name}.conv2.bias"], padding=1))
    sc = q(F.conv2d(x, q(params[f"{name}.shortcut_conv.weight"]), params[f"{name}.shortcut_conv.bias"]))
    return q(F.relu(_bn(sc, params, buffers, f"{name}.shortcut_bn", training) + _bn(c2, params, buffers, f"{name}.bn2", training)))


def soft_argmax(hm: torch.Tensor) -> torch.Tensor:
    """RektNet/keypoint_net.py:51-56 -- expectation of x and y under the heat-map, coordinates i/N."""
    H, W = hm.shape[-2], hm.shape[-1]
    vy = torch.linspace(0, (H - 1.0) / H, H, dtype=hm.dtype)
    vx = torch.linspace(0, (W - 1.0) / W, W, dtype=hm.dtype)
    ey = (hm.sum(3) * vy).sum(-1)
    ex = (hm.sum(2) * vx).sum(-1)
    return torch.stack([ex, ey], -1)


def keypointnet_forward(params, buffers, x, training: bool = True, onnx_mode: bool = False, emulate_bf16: bool = False):
    """RektNet/keypoint_net.py:58-70.  Returns (heat-map [B,K,H,W] softmaxed over H*W, points [B,K,2]).

    emulate_bf16=True is NOT the reference algorithm: it also rounds what the B200 path stores in bf16
    (input, conv weights, conv outputs, activations), to separate kernel errors from storage precision."""
    q = _StoreBf16.apply if emulate_bf16 else _ident
    a = q(F.relu(_bn(q(F.conv2d(q(x), q(params["conv.weight"]), params["conv.bias"], padding=3)), params, buffers, "bn",
                     training)))
    for name, _, _ in BLOCKS:
        a = res_block(a, params, buffers, name, training, q)
    logits = F.conv2d(a, q(params["out.weight"]), params["out.bias"])  # computed twice in the reference (:64,:68)
    if onnx_mode:
        return logits
    B, K, H, W = logits.shape
    hm = F.softmax(logits.view(-1, H * W), 1).view(B, K, H, W)  # :46-49
    return hm, soft_argmax(hm).view(-1, K, 2)


_UNIT = [(5, 3), (3, 1), (1, 0), (6, 4), (4, 2), (2, 0), (2, 1), (4, 3), (6, 5)]  # v53 v31 v10 v64 v42 v20 h21 h43 h65


def cross_ratio_loss(hm, pts, thm, tpts, loss_type: str, include_geo: bool, gamma_horz: float, gamma_vert: float):
    """RektNet/cross_ratio_loss.py:20-63.  Returns (location, geo, total).  The geometric term keeps the
    reference's B x B tensordot (all sample pairs), NOT the per-sample form of the paper."""
    if loss_type in ("l2_softargmax", "l2_sm"):
        loc = ((pts - tpts) ** 2).sum(2).sum(1).mean()
    elif loss_type in ("l2_heatmap", "l2_hm"):
        loc = ((hm - thm) ** 2).sum(3).sum(2).sum(1).mean()
    elif loss_type in ("l1_softargmax", "l1_sm"):
        loc = torch.abs(pts - tpts).sum(2).sum(1).mean()
    else:
        raise ValueError("Did not recognize loss function selection!")
    if include_geo:
        u = {ij: F.normalize(pts[:, ij[0]] - pts[:, ij[1]], dim=1) for ij in _UNIT}
        pair = lambda a, b: 1.0 - torch.tensordot(u[a], u[b], dims=([1], [1]))  # [B,B]
        vA, vB = pair((3, 1), (5, 3)), pair((1, 0), (3, 1))
        vC, vD = pair((6, 4), (4, 2)), pair((4, 2), (2, 0))
        hA, hB = pair((4, 3), (2, 1)), pair((6, 5), (4, 3))
        geo = gamma_horz * (hA + hB).mean() / 2 + gamma_vert * (vA + vB + vC + vD).mean() / 4
    else:
        geo = torch.tensor(0)
    return loc, geo, loc + geo


# --------------------------------------------------------------------------- synthetic inputs
def synth_batch(B: int, seed: int = 0, size: int = 80, num_kpt: int = 7):
    """Images U[0,1); cone-like target points; target heat-maps = delta -> 5x5 Gaussian (sigma 1.1, the
    cv2.GaussianBlur((5,5),0) recipe of RektNet/utils.py:83-96) -> normalised to sum 1."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, 3, size, size, generator=g)
    top = torch.stack([0.4 + 0.2 * torch.rand(B, generator=g), 0.1 + 0.1 * torch.rand(B, generator=g)], -1)
    pts = torch.zeros(B, num_kpt, 2)
    pts[:, 0] = top
    for level in range(3):
        y = top[:, 1] + (level + 1) * (0.2 + 0.05 * torch.rand(B, generator=g))
        half = (level + 1) * (0.08 + 0.03 * torch.rand(B, generator=g))
        pts[:, 1 + 2 * level] = torch.stack([top[:, 0] - half, y], -1)
        pts[:, 2 + 2 * level] = torch.stack([top[:, 0] + half, y], -1)
    pts = pts.clamp(0.1, 0.9)
    k1 = torch.tensor([math.exp(-((i - 2) ** 2) / (2 * 1.1 ** 2)) for i in range(5)])
    k1 = k1 / k1.sum()
    k2 = (k1[:, None] * k1[None, :]).view(1, 1, 5, 5)
    hm = torch.zeros(B * num_kpt, 1, size, size)
    ix = (pts[..., 0] * size).long().clamp(0, size - 1).view(-1)
    iy = (pts[..., 1] * size).long().clamp(0, size - 1).view(-1)
    hm[torch.arange(B * num_kpt), 0, iy, ix] = 1.0
    hm = F.conv2d(hm, k2, padding=2)
    hm = hm / hm.sum(dim=(2, 3), keepdim=True)
    return x, hm.view(B, num_kpt, size, size), pts
