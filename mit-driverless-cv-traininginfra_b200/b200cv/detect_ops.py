"""Detection post-processing on top of the C ABI: confidence filter + greedy NMS
(CVC-YOLOv3/detect.py:84-90, utils/nms.py:4-61) and the crop + cv2-style resize that feeds RektNet
(RektNet/utils.py:73-76, RektNet/detect.py:32-34)."""
from __future__ import annotations

from dataclasses import dataclass

import torch

from .lib import lib, ptr, require_cuda, stream_ptr

MAX_TOP_K = 512


@dataclass
class Detections:
    boxes: torch.Tensor   # fp32 [B, top_k, 4]  corners (x1,y1,x2,y2) in network-input pixels, visiting order
    scores: torch.Tensor  # fp32 [B, top_k]
    rows: torch.Tensor    # int32 [B, top_k]    row of the Darknet eval output (-1 past counts[b])
    counts: torch.Tensor  # int32 [B]
    top_k: int


def detect_nms(det: torch.Tensor, conf_thres: float, nms_thres: float, top_k: int = 200,
               corners: bool = False) -> Detections:
    """det = Darknet eval output [B, rows, 5+C] (or [rows, 5+C]); one CTA per image, no host synchronisation.
    corners=True: rows are (x1,y1,x2,y2,score,...) instead of (cx,cy,w,h,conf,...)."""
    require_cuda(det, "detect_nms")
    if det.dim() == 2:
        det = det.unsqueeze(0)
    if det.dim() != 3 or det.shape[-1] < 5:
        raise ValueError(f"detect_nms: expected [B, rows, >=5], got {tuple(det.shape)}")
    det = det.float()
    if det.stride(2) != 1 or det.stride(1) != det.shape[2]:
        det = det.contiguous()
    b, rows, rl = det.shape
    dev = det.device
    out = Detections(torch.empty(b, top_k, 4, dtype=torch.float32, device=dev),
                     torch.empty(b, top_k, dtype=torch.float32, device=dev),
                     torch.empty(b, top_k, dtype=torch.int32, device=dev),
                     torch.empty(b, dtype=torch.int32, device=dev), top_k)
    if not 1 <= top_k <= MAX_TOP_K:
        raise ValueError(f"detect_nms: top_k={top_k} outside [1,{MAX_TOP_K}]")
    lib().call("b200cv_detect_nms", ptr(det), det.stride(0), b, rows, rl, int(corners), float(conf_thres),
               float(nms_thres),
               int(top_k), ptr(out.boxes), ptr(out.scores), ptr(out.rows), ptr(out.counts), stream_ptr())
    return out


def compact(d: Detections):
    """counts -> (offsets int32 [B+1], src int32 [B*top_k, 2] = (image, slot) per crop)."""
    b = d.counts.shape[0]
    if b == 0:
        return (torch.zeros(1, dtype=torch.int32, device=d.counts.device),
                torch.zeros(0, 2, dtype=torch.int32, device=d.counts.device))
    offsets = torch.empty(b + 1, dtype=torch.int32, device=d.counts.device)
    src = torch.empty(b * d.top_k, 2, dtype=torch.int32, device=d.counts.device)
    lib().call("b200cv_detect_compact", ptr(d.counts), b, d.top_k, ptr(offsets), ptr(src), stream_ptr())
    return offsets, src


def crop_resize(frames: torch.Tensor, d: Detections, src: torch.Tensor, n_crops: int, geom: torch.Tensor,
                size=(80, 80)):
    """frames u8 [B,H,W,3] (BGR, as cv2.imread gives them); geom fp32 [3] (shared) or [B,3] = (ratio, pad_w, pad_h).
    Returns (crops fp32 [n,3,h,w] ready for KeypointNet, rects int32 [n,4])."""
    require_cuda(frames, "crop_resize")
    if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[-1] != 3 or not frames.is_contiguous():
        raise ValueError("crop_resize: frames must be a contiguous uint8 [B,H,W,3] tensor")
    b, h, w, _ = frames.shape
    if b != d.counts.shape[0]:
        raise ValueError("crop_resize: frames and detections disagree on the batch size")
    geom = geom.to(device=frames.device, dtype=torch.float32).contiguous()
    gstride = 0 if geom.dim() == 1 else geom.shape[1]
    if geom.dim() == 2 and geom.shape[0] != b:
        raise ValueError("crop_resize: geom must be [3] or [B,3]")
    ow, oh = int(size[0]), int(size[1])
    out = torch.empty(n_crops, 3, oh, ow, dtype=torch.float32, device=frames.device)
    rects = torch.empty(n_crops, 4, dtype=torch.int32, device=frames.device)
    lib().call("b200cv_crop_resize_u8", ptr(frames), b, h, w, ptr(d.boxes), d.top_k, ptr(src), int(n_crops),
               ptr(geom), gstride, ow, oh, ptr(out), ptr(rects), stream_ptr())
    return out, rects


@dataclass
class ImageMetrics:
    ap: torch.Tensor         # fp32 [B]  average precision per image (0 where skipped)
    recall: torch.Tensor     # fp32 [B]
    precision: torch.Tensor  # fp32 [B]
    valid: torch.Tensor      # int32 [B] 0 = the reference skips the image (no detection / no label)
    correct: torch.Tensor    # uint8 [B, top_k] true-positive flag of each kept detection

    def means(self):
        """(mean AP, mean recall, mean precision) over the images the reference counts (validate.py:170-174)."""
        m = self.valid.bool()
        if not bool(m.any()):
            return float("nan"), float("nan"), float("nan")
        return float(self.ap[m].mean()), float(self.recall[m].mean()), float(self.precision[m].mean())


def match_ap(d: Detections, targets: torch.Tensor, width: float, height: float, iou_thres: float) -> ImageMetrics:
    """Per-image AP / recall / precision of the NMS output against normalised labels [B,T,5] (validate.py:98-130,
    utils/utils.py:58-119), single class like the reference."""
    require_cuda(targets, "match_ap")
    targets = targets.float().contiguous()
    b = d.counts.shape[0]
    if targets.dim() != 3 or targets.shape[0] != b or targets.shape[2] != 5:
        raise ValueError(f"match_ap: expected targets [B={b}, T, 5], got {tuple(targets.shape)}")
    dev = targets.device
    f = lambda: torch.empty(b, dtype=torch.float32, device=dev)
    out = ImageMetrics(f(), f(), f(), torch.empty(b, dtype=torch.int32, device=dev),
                       torch.empty(b, d.top_k, dtype=torch.uint8, device=dev))
    lib().call("b200cv_detect_match_ap", ptr(d.boxes), ptr(d.counts), b, d.top_k, ptr(targets), targets.shape[1],
               float(width), float(height), float(iou_thres), ptr(out.ap), ptr(out.recall), ptr(out.precision),
               ptr(out.valid), ptr(out.correct), stream_ptr())
    return out
