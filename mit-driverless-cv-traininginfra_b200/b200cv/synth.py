"""Synthetic inputs of the named shapes for bench.py and the measurement tools: there is no network for
datasets, so images, cone-like targets, keypoint batches and camera frames are generated from seeds.  Plain
torch / numpy on the host -- no kernels, no oracle.  (The oracle keeps its own copies of the same recipes for the
parity tests; tests/test_host_logic.py checks that both produce identical tensors.)"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


def synth_images(B: int, H: int, W: int, seed: int = 0) -> torch.Tensor:
    """fp32 [B,3,H,W] in [0,1) -- the range `to_tensor` produces (CVC-YOLOv3/utils/datasets.py:311)."""
    return torch.rand(B, 3, H, W, generator=torch.Generator().manual_seed(seed))


def synth_targets(B: int, T: int = 16, seed: int = 1) -> torch.Tensor:
    """Cone-like boxes [B,T,5]: n in [1,T] per image, (cls=0, cx, cy, w, h) normalised, remaining rows zero."""
    g = torch.Generator().manual_seed(seed)
    t = torch.zeros(B, T, 5)
    for b in range(B):
        n = int(torch.randint(1, T + 1, (1,), generator=g))
        t[b, :n, 1:3] = 0.05 + 0.9 * torch.rand(n, 2, generator=g)
        t[b, :n, 3] = 0.01 + 0.08 * torch.rand(n, generator=g)
        t[b, :n, 4] = 0.02 + 0.16 * torch.rand(n, generator=g)
    return t


def synth_keypoint_batch(B: int, seed: int = 0, size: int = 80, num_kpt: int = 7):
    """RektNet batch: images U[0,1); cone-like target points; target heat-maps = delta -> 5x5 Gaussian (sigma 1.1, the
    cv2.GaussianBlur((5,5),0) recipe of RektNet/utils.py:83-96) -> normalised to sum 1.  Returns (x, heatmaps, points)."""
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


def synth_frames(B: int, H: int, W: int, seed: int = 0) -> np.ndarray:
    """Camera-like u8 frames [B,H,W,3]: smooth gradients + noise."""
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    out = np.empty((B, H, W, 3), np.uint8)
    for b in range(B):
        base = np.stack([(xx * (b + 1) + yy) % 256, (xx + 2 * yy * (b + 1)) % 256, (xx * yy // 7) % 256], -1)
        out[b] = ((base + rng.randint(0, 64, size=(H, W, 3))) % 256).astype(np.uint8)
    return out


def conv_layer_table(model):
    """Layer list of a models.Darknet from its executor, in the form bench.py's FLOP counter walks:
    {"type", "cin", "cout", "k", "stride", "pad", "layers"} (route inputs as absolute layer indices)."""
    rows = []
    for L in model.engine().layers:
        row = {"type": L.type, "layers": list(L.inputs)}
        if L.type == "convolutional":
            row.update(cin=L.cin, cout=L.cout, k=L.k, stride=L.stride, pad=L.pad)
        elif L.type == "maxpool":
            row["stride"] = L.pool_stride
        rows.append(row)
    return rows
