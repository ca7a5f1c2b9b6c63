"""YOLO head operators on top of the C ABI: target assignment, loss (+ gradient of the raw head),
eval decode, and the autograd.Function behind the public ``models.YOLOLayer.forward``."""
from __future__ import annotations

from dataclasses import dataclass

import torch

from .lib import DT_BF16, DT_F32, lib, ptr, require_cuda, stream_ptr


@dataclass
class YoloTargets:
    owner: torch.Tensor   # int32 [B,A,Gh,Gw]  winning target index or -1
    ign: torch.Tensor     # uint8 [Gh,Gw]      cross-batch ignore cells
    rec: torch.Tensor     # fp32  [B,T,8]      per-target record
    counts: torch.Tensor  # int32 [2]          (N_mask, N_conf_false)
    B: int
    T: int
    A: int
    Gh: int
    Gw: int


def scaled_anchors(anchors, stride: float, device) -> torch.Tensor:
    # divided in Python float64, then cast to fp32 -- CVC-YOLOv3/models.py:160
    return torch.tensor([(a_w / stride, a_h / stride) for a_w, a_h in anchors], dtype=torch.float, device=device)


def yolo_targets(targets: torch.Tensor, anchors_scaled: torch.Tensor, gh: int, gw: int, ignore_thres: float) -> YoloTargets:
    require_cuda(targets, "yolo_targets")
    targets = targets.contiguous().float()
    b, t, _ = targets.shape
    a = anchors_scaled.shape[0]
    dev = targets.device
    yt = YoloTargets(torch.empty(b, a, gh, gw, dtype=torch.int32, device=dev),
                     torch.empty(gh, gw, dtype=torch.uint8, device=dev),
                     torch.empty(b, t, 8, dtype=torch.float32, device=dev),
                     torch.empty(2, dtype=torch.int32, device=dev), b, t, a, gh, gw)
    lib().call("b200cv_yolo_targets", ptr(targets), ptr(anchors_scaled), b, t, a, gh, gw, float(ignore_thres),
               ptr(yt.owner), ptr(yt.ign), ptr(yt.rec), ptr(yt.counts), stream_ptr())
    return yt


def yolo_targets_dense(yt: YoloTargets, num_classes: int):
    """The eight tensors of utils.utils.build_targets (u8, u8, f32 x5, u8)."""
    dev = yt.owner.device
    shape = (yt.B, yt.A, yt.Gh, yt.Gw)
    mask = torch.empty(shape, dtype=torch.uint8, device=dev)
    conf_mask = torch.empty(shape, dtype=torch.uint8, device=dev)
    f = [torch.empty(shape, dtype=torch.float32, device=dev) for _ in range(5)]
    tcls = torch.empty(shape + (num_classes,), dtype=torch.uint8, device=dev)
    lib().call("b200cv_yolo_targets_dense", ptr(yt.owner), ptr(yt.ign), ptr(yt.rec), yt.B, yt.T, yt.A, num_classes,
               yt.Gh, yt.Gw, ptr(mask), ptr(conf_mask), ptr(f[0]), ptr(f[1]), ptr(f[2]), ptr(f[3]), ptr(f[4]),
               ptr(tcls), stream_ptr())
    return mask, conf_mask, f[0], f[1], f[2], f[3], f[4], tcls


def _head_strides(z: torch.Tensor, nchw: bool):
    """(sb, sy, sx, sc) element strides of a head tensor given as NCHW [B,ch,G,G] or NHWC [B,G,G,ch]."""
    if nchw:
        return z.stride(0), z.stride(2), z.stride(3), z.stride(1)
    return z.stride(0), z.stride(1), z.stride(2), z.stride(3)


def yolo_loss(z, nchw, yt: YoloTargets, num_classes, consts, sums=None, dlogits=None, gscale=None):
    """consts = (xy, wh, obj, noobj).  sums: fp64[6] accumulated; dlogits: same layout family as z."""
    xy, wh, obj, noobj = consts
    zs = _head_strides(z, nchw)
    if dlogits is not None:
        ds = _head_strides(dlogits, nchw)
        dch = dlogits.shape[1] if nchw else dlogits.shape[-1]
        ddt = DT_F32 if dlogits.dtype == torch.float32 else DT_BF16
    else:
        ds, dch, ddt = (0, 0, 0, 0), 0, DT_BF16
    lib().call("b200cv_yolo_loss", ptr(z), *zs, yt.B, yt.A, num_classes, yt.Gh, yt.Gw, ptr(yt.owner), ptr(yt.ign),
               ptr(yt.rec), yt.T, ptr(yt.counts), float(xy), float(wh), float(obj), float(noobj), ptr(sums),
               ptr(dlogits), ddt, *ds, dch, ptr(gscale), stream_ptr())


def yolo_loss_cells(z, nchw, yt: YoloTargets, num_classes, consts, sums=None, dcell=None, gscale=None):
    """One thread per anchor cell: loss sums and/or compact cell gradients dcell [B*Gh*Gw, >=5A] fp32."""
    xy, wh, obj, noobj = consts
    lib().call("b200cv_yolo_loss_cells", ptr(z), *_head_strides(z, nchw), yt.B, yt.A, num_classes, yt.Gh, yt.Gw,
               ptr(yt.owner), ptr(yt.ign), ptr(yt.rec), yt.T, ptr(yt.counts), float(xy), float(wh), float(obj),
               float(noobj), ptr(sums), ptr(dcell), 0 if dcell is None else dcell.shape[-1], ptr(gscale), stream_ptr())


def yolo_expand_dlogits(dcell, dlogits, num_anchors, num_classes):
    """Compact cell gradients -> dense NHWC head gradient [B,G,G,Cpad] (zeros for class / pad channels)."""
    npix = dlogits.numel() // dlogits.shape[-1]
    lib().call("b200cv_yolo_expand_dlogits", ptr(dcell), dcell.shape[-1], ptr(dlogits),
               DT_F32 if dlogits.dtype == torch.float32 else DT_BF16, dlogits.stride(-2), dlogits.shape[-1], npix,
               num_anchors, num_classes, stream_ptr())


def yolo_head_grad(z, yt: YoloTargets, num_classes, consts, gscale, out_dtype=torch.bfloat16):
    """dlogits (NHWC, same shape as z) of the YOLO loss: cell kernel + streaming expansion."""
    npix = z.shape[0] * z.shape[1] * z.shape[2]
    cell_ld = (5 * yt.A + 3) // 4 * 4
    dcell = torch.empty(npix, cell_ld, dtype=torch.float32, device=z.device)
    yolo_loss_cells(z, False, yt, num_classes, consts, dcell=dcell, gscale=gscale)
    from . import ops

    if ops.split_mode() and out_dtype == torch.bfloat16:  # fp32-parity mode: fp32 gradient, stored split
        dl = torch.empty(z.shape, dtype=torch.float32, device=z.device)
        yolo_expand_dlogits(dcell, dl, yt.A, num_classes)
        return ops.split_from_f32(dl)
    dl = torch.empty(z.shape, dtype=out_dtype, device=z.device)
    yolo_expand_dlogits(dcell, dl, yt.A, num_classes)
    return dl


def yolo_loss_finalize(sums, yt: YoloTargets, consts, out7):
    xy, wh, obj, noobj = consts
    lib().call("b200cv_yolo_loss_finalize", ptr(sums), ptr(yt.counts), float(xy), float(wh), float(obj), float(noobj),
               ptr(out7), stream_ptr())


def yolo_decode(z, nchw, num_anchors, num_classes, anchors_scaled, stride, out, row_offset):
    b = z.shape[0]
    gh, gw = (z.shape[2], z.shape[3]) if nchw else (z.shape[1], z.shape[2])
    lib().call("b200cv_yolo_decode", ptr(z), *_head_strides(z, nchw), b, num_anchors, num_classes, gh, gw,
               ptr(anchors_scaled), float(stride), ptr(out), out.stride(0), int(row_offset), stream_ptr())


class YoloLayerFn(torch.autograd.Function):
    """loss7 = f(sample NCHW fp32, targets): element 0 is the differentiable total, 1..6 the parts."""

    @staticmethod
    def forward(ctx, sample, targets, anchors_scaled, num_classes, ignore_thres, consts):
        require_cuda(sample, "YOLOLayer")
        sample = sample.contiguous().float()
        _, _, gh, gw = sample.shape
        yt = yolo_targets(targets, anchors_scaled, gh, gw, ignore_thres)
        sums = torch.zeros(6, dtype=torch.float64, device=sample.device)
        out7 = torch.zeros(7, dtype=torch.float32, device=sample.device)
        yolo_loss(sample, True, yt, num_classes, consts, sums=sums)
        yolo_loss_finalize(sums, yt, consts, out7)
        ctx.save_for_backward(sample)
        ctx.yt, ctx.num_classes, ctx.consts = yt, num_classes, consts
        return out7

    @staticmethod
    def backward(ctx, g7):
        (sample,) = ctx.saved_tensors
        g = g7[0:1].contiguous().float()
        dsample = torch.zeros_like(sample)
        yolo_loss(sample, True, ctx.yt, ctx.num_classes, ctx.consts, dlogits=dsample, gscale=g)
        return dsample, None, None, None, None, None
