"""detect -> NMS -> crop -> RektNet, the inference joint of BASELINE config 5.

The reference runs the two networks from separate scripts (CVC-YOLOv3/detect.py:60-98 draws boxes,
RektNet/detect.py:20-58 takes an already-cropped cone image); nothing in it joins them.  This module defines the
joint from the reference's own pieces, entirely on the device:

    det   = darknet(imgs)                       models.Darknet.forward, eval mode          (models.py:312-338)
    boxes = conf filter + corners + greedy NMS  detect.py:84-90, utils/nms.py:4-61         (b200cv_detect_nms)
    crops = cv2.resize(frame[rect], (80,80))    detect.py:93-96 box -> frame mapping,
            .transpose(2,0,1) / 255.0           RektNet/utils.py:73-76, RektNet/detect.py:32-34 (b200cv_crop_resize_u8)
    hm, pts = keypoint_net(crops)               keypoint_net.KeypointNet.forward, eval mode

One 4-byte device->host read (the number of crops) sits between NMS and the crop launch; everything else is
stream-ordered.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

from . import detect_ops
from .lib import require_cuda


@dataclass
class PipelineOutput:
    detections: detect_ops.Detections  # boxes / scores / rows / counts per image (network-input pixels)
    offsets: torch.Tensor              # int32 [B+1]: crops of image b are [offsets[b], offsets[b+1])
    rects: torch.Tensor                # int32 [n,4]: crop rectangle (x0,y0,x1,y1) in frame pixels
    points: torch.Tensor               # fp32 [n,7,2]: keypoints, normalised to the crop (keypoint_net.py:51-56)
    heatmaps: Optional[torch.Tensor]   # fp32 [n,7,80,80] when keep_heatmaps
    n_crops: int

    def points_in_frame(self) -> torch.Tensor:
        """Keypoints in frame pixels: x = x0 + px * crop_w (the scaling of RektNet/utils.py:64)."""
        r = self.rects.float()
        wh = torch.stack([r[:, 2] - r[:, 0], r[:, 3] - r[:, 1]], 1)
        return r[:, None, 0:2] + self.points * wh[:, None, :]


class ConePipeline:
    def __init__(self, darknet, keypoint_net, conf_thres: Optional[float] = None, nms_thres: Optional[float] = None,
                 top_k: int = 200, kpt_size=(80, 80), max_crops_per_pass: int = 1024):
        self.darknet = darknet
        self.keypoint_net = keypoint_net
        cfg_conf, cfg_nms, _ = darknet.get_threshs()  # the cfg's conf_thresh / nms_thresh (yolo_baseline.cfg:18-19)
        self.conf_thres = float(cfg_conf if conf_thres is None else conf_thres)
        self.nms_thres = float(cfg_nms if nms_thres is None else nms_thres)
        self.top_k = int(top_k)
        self.kpt_size = tuple(kpt_size)
        self.max_crops_per_pass = int(max_crops_per_pass)
        self._letterbox = {}

    @torch.no_grad()
    def from_frames(self, frames: torch.Tensor, keep_heatmaps: bool = False) -> PipelineOutput:
        """The joint from raw camera frames: u8 [B,H,W,3] BGR (cv2 order) on the device.  The network input is made by
        the letterbox kernel (detect.py:62-72: pad 127, PIL-bilinear resize, /255; BGR -> RGB planes), the crops are
        cut from the same frames."""
        require_cuda(frames, "ConePipeline.from_frames")
        key = (tuple(frames.shape[1:3]), str(frames.device))
        if key not in self._letterbox:
            from .preprocess import Letterbox

            self._letterbox[key] = Letterbox(frames.shape[1:3], self.darknet.img_size(), frames.device)
        lb = self._letterbox[key]
        return self(lb(frames, reverse_channels=True), frames, lb.geom, keep_heatmaps)

    @torch.no_grad()
    def __call__(self, imgs: torch.Tensor, frames: torch.Tensor, geom: torch.Tensor,
                 keep_heatmaps: bool = False) -> PipelineOutput:
        """imgs fp32 [B,3,S,S] network input; frames u8 [B,H,W,3] (BGR) the crops are cut from;
        geom fp32 [3] or [B,3] = (ratio, pad_w, pad_h) of the letterbox that made imgs from frames."""
        require_cuda(imgs, "ConePipeline")
        require_cuda(frames, "ConePipeline")
        if self.darknet.training or self.keypoint_net.training:
            raise RuntimeError("ConePipeline: put both networks in eval() mode")
        det = self.darknet(imgs)
        d = detect_ops.detect_nms(det, self.conf_thres, self.nms_thres, self.top_k)
        offsets, src = detect_ops.compact(d)
        n = int(offsets[-1])  # the one host synchronisation of the pipeline
        crops, rects = detect_ops.crop_resize(frames, d, src, n, geom, self.kpt_size)
        pts, hms = [], []
        for i in range(0, n, self.max_crops_per_pass):
            hm, p = self.keypoint_net(crops[i:i + self.max_crops_per_pass])
            pts.append(p)
            if keep_heatmaps:
                hms.append(hm)
        k = getattr(self.keypoint_net, "num_kpt", 7)
        points = torch.cat(pts, 0) if pts else torch.empty(0, k, 2, device=imgs.device)
        heat = (torch.cat(hms, 0) if hms else torch.empty(0, k, self.kpt_size[1], self.kpt_size[0],
                                                          device=imgs.device)) if keep_heatmaps else None
        return PipelineOutput(d, offsets, rects, points, heat, n)


@torch.no_grad()
def detection_metrics(darknet, imgs: torch.Tensor, targets: torch.Tensor, conf_thres: Optional[float] = None,
                      nms_thres: Optional[float] = None, iou_thres: Optional[float] = None,
                      top_k: int = 200) -> detect_ops.ImageMetrics:
    """The per-batch body of CVC-YOLOv3/validate.py:73-130 on the device: eval forward, confidence filter + NMS, greedy
    matching against the labels, AP / recall / precision per image -- three launches after the network, no per-image
    Python loop and no host synchronisation.  Thresholds default to the cfg's (validate.py:66).  `.means()` of the
    result is what validate() averages over the images of a batch."""
    require_cuda(imgs, "detection_metrics")
    if darknet.training:
        raise RuntimeError("detection_metrics: put the network in eval() mode (validate.py:68)")
    cfg_conf, cfg_nms, cfg_iou = darknet.get_threshs()
    width, height = darknet.img_size()
    det = darknet(imgs)
    d = detect_ops.detect_nms(det, cfg_conf if conf_thres is None else conf_thres,
                              cfg_nms if nms_thres is None else nms_thres, top_k)
    return detect_ops.match_ap(d, targets.to(imgs.device), width, height, cfg_iou if iou_thres is None else iou_thres)
