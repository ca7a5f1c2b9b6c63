"""Tile-and-scale input pipeline on the device: the `ts` branch of CVC-YOLOv3/utils/datasets.py:143-159
for a batch of camera frames in one kernel launch, bit-exact with Pillow:

    scaled = scale_image(img, scale)                        utils/utils.py:321-326   (PIL resize, ANTIALIAS = LANCZOS)
    vert_pad, horiz_pad = pre_tile_padding(...)             utils/utils.py:376-382
    padded = pad(scaled, fill=127)                          torchvision.transforms.functional.pad
    patch, boundary = get_patch(padded, w, h, patch_index)  utils/utils.py:411-426   (PIL crop rounds its float box)
    img = to_tensor(patch)                                  CHW float32 / 255

and the matching label transform (datasets.py:176-186, utils/utils.py:456-472, datasets.py:300-313) on the host --
a handful of boxes per frame.  The LANCZOS coefficient tables depend only on the geometry: they are computed here on the
host exactly as Pillow's precompute_coeffs / normalize_coeffs_8bpc do and cached on the device; neither the scaled nor
the padded frame is materialised.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .lib import lib, ptr, require_cuda, stream_ptr
from .preprocess import PRECISION_BITS


def lanczos_tables(in_size: int, out_size: int):
    """Pillow's LANCZOS resampling tables for one axis: (first source index [out], tap count [out], 22-bit fixed-point
    coefficients [out, ksize]) as int32 arrays (Resample.c: lanczos_filter, support 3, widened by the scale factor
    when the axis shrinks)."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 3.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    inv = 1.0 / filterscale
    center = (np.arange(out_size, dtype=np.float64) + 0.5) * scale
    xmin = np.maximum(np.trunc(center - support + 0.5).astype(np.int64), 0)
    xmax = np.minimum(np.trunc(center + support + 0.5).astype(np.int64), in_size)
    cnt = xmax - xmin
    taps = np.arange(ksize, dtype=np.int64)[None, :]
    arg = ((taps + xmin[:, None]).astype(np.float64) - center[:, None] + 0.5) * inv

    def sinc(x):
        px = x * math.pi
        with np.errstate(invalid="ignore", divide="ignore"):
            return np.where(x == 0.0, 1.0, np.sin(px) / np.where(px == 0.0, 1.0, px))

    w = np.where((arg >= -3.0) & (arg < 3.0), sinc(arg) * sinc(arg / 3), 0.0)
    w = np.where(taps < cnt[:, None], w, 0.0)
    ww = np.zeros(out_size, np.float64)
    for x in range(ksize):  # sequential sum, the order of Pillow's loop
        ww = ww + w[:, x]
    k = np.where(ww[:, None] != 0.0, w / np.where(ww == 0.0, 1.0, ww)[:, None], w)
    v = k * float(1 << PRECISION_BITS)
    kk = np.where(k < 0, np.trunc(-0.5 + v), np.trunc(0.5 + v)).astype(np.int32)
    return xmin.astype(np.int32), cnt.astype(np.int32), kk


def pre_tile_padding(img_width, img_height, patch_width, patch_height):
    vert_pad = math.ceil((patch_height - img_height) / 2) if img_height < patch_height else 0
    horiz_pad = math.ceil((patch_width - img_width) / 2) if img_width < patch_width else 0
    return vert_pad, horiz_pad


def get_patch_spacings(img_width, img_height, patch_width, patch_height):
    """(patches wide, patches high, total, horizontal overlap step, vertical overlap step) -- utils/utils.py:384-405."""
    if img_width < patch_width or img_height < patch_height:
        raise ValueError("image smaller than the patch")
    hn = math.ceil(img_width / patch_width)
    h_off = 0 if hn == 1 else (hn * patch_width - img_width) / (hn - 1)
    vn = math.ceil(img_height / patch_height)
    v_off = 0 if vn == 1 else (vn * patch_height - img_height) / (vn - 1)
    return hn, vn, hn * vn, h_off, v_off


class TileScale:
    """frames u8 RGB [B,H,W,3] on the device + one patch index per frame -> network input fp32 [B,3,patch_h,patch_w]."""

    def __init__(self, frame_hw, scale: float, patch_wh, device, fill: int = 127):
        self.h, self.w = int(frame_hw[0]), int(frame_hw[1])
        self.scale = float(scale)
        self.patch_w, self.patch_h = int(patch_wh[0]), int(patch_wh[1])
        self.new_h, self.new_w = int(self.h * self.scale), int(self.w * self.scale)  # scale_image
        self.vert_pad, self.horiz_pad = pre_tile_padding(self.new_w, self.new_h, self.patch_w, self.patch_h)
        self.padded_w, self.padded_h = self.new_w + 2 * self.horiz_pad, self.new_h + 2 * self.vert_pad
        self.n_wide, self.n_high, self.n_patches, self.h_off, self.v_off = get_patch_spacings(
            self.padded_w, self.padded_h, self.patch_w, self.patch_h)
        self.fill = int(fill)
        self.device = torch.device(device)
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.device)
        self.hx = tuple(dev(a) for a in lanczos_tables(self.w, self.new_w)) if self.w != self.new_w else None
        self.vy = tuple(dev(a) for a in lanczos_tables(self.h, self.new_h)) if self.h != self.new_h else None

    def boundary(self, patch_index: int):
        """(left, top, right, bottom) of get_patch in the padded scaled frame (floats, as the reference returns them)."""
        row_position = patch_index % self.n_wide
        left = self.patch_w * row_position - self.h_off * row_position
        col_position = math.floor(patch_index / self.n_wide)
        top = self.patch_h * col_position - self.v_off * col_position
        return (left, top, left + self.patch_w, top + self.patch_h)

    def __call__(self, frames: torch.Tensor, patch_index) -> torch.Tensor:
        require_cuda(frames, "TileScale")
        if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[-1] != 3 or not frames.is_contiguous():
            raise ValueError("TileScale: frames must be a contiguous uint8 [B,H,W,3] tensor")
        if tuple(frames.shape[1:3]) != (self.h, self.w):
            raise ValueError(f"TileScale: built for {self.h}x{self.w} frames, got {tuple(frames.shape[1:3])}")
        b = frames.shape[0]
        idx = [int(patch_index)] * b if isinstance(patch_index, int) else [int(i) for i in patch_index]
        if len(idx) != b or any(i < 0 or i >= self.n_patches for i in idx):
            raise ValueError("TileScale: one patch index in [0, n_patches) per frame")
        bounds = [self.boundary(i) for i in idx]
        # PIL.Image.crop rounds its box (Python round: half to even); subtract the pad to get scaled-frame coordinates
        offs = torch.tensor([[int(round(l)) - self.horiz_pad for l, _, _, _ in bounds],
                             [int(round(t)) - self.vert_pad for _, t, _, _ in bounds]], dtype=torch.int32)
        offs = offs.to(frames.device)
        out = torch.empty(b, 3, self.patch_h, self.patch_w, dtype=torch.float32, device=frames.device)
        hx = self.hx or (None, None, None)
        vy = self.vy or (None, None, None)
        lib().call("b200cv_tile_scale_u8", ptr(frames), b, self.h, self.w, self.new_w, self.new_h, self.fill,
                   ptr(offs[0]), ptr(offs[1]), ptr(hx[0]), ptr(hx[1]), ptr(hx[2]),
                   0 if hx[2] is None else hx[2].shape[1], ptr(vy[0]), ptr(vy[1]), ptr(vy[2]),
                   0 if vy[2] is None else vy[2].shape[1], self.patch_w, self.patch_h, ptr(out), stream_ptr())
        return out

    def labels(self, labels_xyhw, patch_index: int, num_targets: int) -> torch.Tensor:
        """[num_targets, 5] rows (0, cx, cy, w, h) normalised by the patch size, zero rows = padding: the label chain of
        datasets.py:176-186,300-313 for one frame.  labels_xyhw: rows (x, y, h, w), (x, y) = upper-left corner."""
        left, top, right, bottom = self.boundary(patch_index)
        out = torch.zeros(num_targets, 5)
        n = 0
        for x, y, hh, ww in (tuple(float(v) for v in row) for row in labels_xyhw):
            x0, y0 = self.scale * x + self.horiz_pad, self.scale * y + self.vert_pad
            x1, y1 = self.scale * (x + ww) + self.horiz_pad, self.scale * (y + hh) + self.vert_pad
            area = (x1 - x0) * (y1 - y0)
            dx, dy = min(x1, right) - max(x0, left), min(y1, bottom) - max(y0, top)
            overlap = dx * dy if (dx >= 0 and dy >= 0) else 0.0
            if area > 0 and (overlap / area > 0.5 or overlap > 1000) and n < num_targets:
                nx0, ny0, nx1, ny1 = max(x0, left) - left, max(y0, top) - top, min(x1, right) - left, min(y1, bottom) - top
                out[n] = torch.tensor([0.0, (nx0 + nx1) / 2 / self.patch_w, (ny0 + ny1) / 2 / self.patch_h,
                                       abs(nx1 - nx0) / self.patch_w, abs(ny1 - ny0) / self.patch_h])
                n += 1
        return out
