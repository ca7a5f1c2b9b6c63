"""Letterbox front end of the detector on the device: the `pad -> resize -> to_tensor` chain of
CVC-YOLOv3/detect.py:62-72 (also validate.py and the scale step of utils/datasets.py) in one kernel launch per batch of
frames, bit-exact with torchvision / Pillow:

    pad_h, pad_w, ratio = calculate_padding(h, w, new_h, new_w)          utils/utils.py:36-48
    img = pad(img, (pad_w, pad_h, pad_w, pad_h), fill=127)               torchvision.transforms.functional.pad
    img = resize(img, (new_h, new_w))                                     PIL BILINEAR (8-bit two-pass resampler)
    img = to_tensor(img)                                                  CHW float32 / 255

The resampling coefficients depend only on the geometry: they are computed here on the host exactly as Pillow's
precompute_coeffs / normalize_coeffs_8bpc do (double precision, same operation order) and cached on the device.
"""
from __future__ import annotations

import numpy as np
import torch

from .lib import lib, ptr, require_cuda, stream_ptr

PRECISION_BITS = 32 - 8 - 2  # Pillow's fixed-point precision for 8-bit images


def calculate_padding(orig_height, orig_width, new_height, new_width):
    """(pad_h, pad_w, scale_factor) -- same arithmetic as the reference's utils.utils.calculate_padding."""
    if max(orig_height, orig_width) == orig_height:
        new_img_width = orig_height * new_width / new_height
        return 0, int((new_img_width - orig_width) / 2), new_height / orig_height
    new_img_height = orig_width * new_height / new_width
    return int((new_img_height - orig_height) / 2), 0, new_width / orig_width


def bilinear_tables(in_size: int, out_size: int):
    """Pillow's BILINEAR resampling tables for one axis: (first source index [out], tap count [out], 22-bit fixed-point
    coefficients [out, ksize]) as int32 arrays.  When it shrinks the axis Pillow widens the triangle filter by the scale
    factor (an anti-aliased reduction), so ksize grows with in_size / out_size."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    inv = 1.0 / filterscale
    center = (np.arange(out_size, dtype=np.float64) + 0.5) * scale
    xmin = np.maximum(np.trunc(center - support + 0.5).astype(np.int64), 0)
    xmax = np.minimum(np.trunc(center + support + 0.5).astype(np.int64), in_size)
    cnt = xmax - xmin
    taps = np.arange(ksize, dtype=np.int64)[None, :]
    arg = np.abs(((taps + xmin[:, None]).astype(np.float64) - center[:, None] + 0.5) * inv)
    w = np.where(arg < 1.0, 1.0 - arg, 0.0)
    w = np.where(taps < cnt[:, None], w, 0.0)
    ww = np.zeros(out_size, np.float64)
    for x in range(ksize):  # sequential sum, the order of Pillow's loop
        ww = ww + w[:, x]
    k = np.where(ww[:, None] != 0.0, w / np.where(ww == 0.0, 1.0, ww)[:, None], w)
    kk = np.trunc(0.5 + k * float(1 << PRECISION_BITS)).astype(np.int32)
    return xmin.astype(np.int32), cnt.astype(np.int32), kk


class Letterbox:
    """frames u8 [B,H,W,3] on the device -> network input fp32 [B,3,new_h,new_w]; `geom` = (ratio, pad_w, pad_h) is what
    maps detections back onto the frame (detect.py:93-96, b200cv.pipeline)."""

    def __init__(self, frame_hw, net_wh, device, fill: int = 127):
        self.h, self.w = int(frame_hw[0]), int(frame_hw[1])
        self.out_w, self.out_h = int(net_wh[0]), int(net_wh[1])
        self.pad_h, self.pad_w, self.ratio = calculate_padding(self.h, self.w, self.out_h, self.out_w)
        self.fill = int(fill)
        self.device = torch.device(device)
        pw, ph = self.w + 2 * self.pad_w, self.h + 2 * self.pad_h
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.device)
        self.hx = tuple(dev(a) for a in bilinear_tables(pw, self.out_w)) if pw != self.out_w else None
        self.vy = tuple(dev(a) for a in bilinear_tables(ph, self.out_h)) if ph != self.out_h else None
        self.geom = torch.tensor([self.ratio, float(self.pad_w), float(self.pad_h)], dtype=torch.float32,
                                 device=self.device)

    def __call__(self, frames: torch.Tensor, reverse_channels: bool = False) -> torch.Tensor:
        """reverse_channels=True: frames are BGR (cv2.imread), the network wants RGB planes (PIL, detect.py:62)."""
        require_cuda(frames, "Letterbox")
        if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[-1] != 3 or not frames.is_contiguous():
            raise ValueError("Letterbox: frames must be a contiguous uint8 [B,H,W,3] tensor")
        if tuple(frames.shape[1:3]) != (self.h, self.w):
            raise ValueError(f"Letterbox: built for {self.h}x{self.w} frames, got {tuple(frames.shape[1:3])}")
        b = frames.shape[0]
        out = torch.empty(b, 3, self.out_h, self.out_w, dtype=torch.float32, device=frames.device)
        hx = self.hx or (None, None, None)
        vy = self.vy or (None, None, None)
        lib().call("b200cv_letterbox_u8", ptr(frames), b, self.h, self.w, self.pad_w, self.pad_h, self.fill,
                   int(reverse_channels), ptr(hx[0]), ptr(hx[1]), ptr(hx[2]), 0 if hx[2] is None else hx[2].shape[1],
                   ptr(vy[0]), ptr(vy[1]), ptr(vy[2]), 0 if vy[2] is None else vy[2].shape[1], self.out_w, self.out_h,
                   ptr(out), stream_ptr())
        return out
