"""Hot-path functions of the reference's utils/utils.py on the B200 kernels.

* ``build_targets``  -- one target-assignment kernel + a dense expansion kernel (reference :195-275)
* ``bbox_iou``       -- element-wise IoU helper (reference :163-193); host utility, any device
* ``weights_init_normal`` -- the init recipe of reference :50-56

The remaining helpers of the reference module (Logger, padding/tiling, AP ...) are not part of the
accelerated path; if B200CV_REFERENCE_ROOT is set they are re-exported from the reference file so
scripts that ``from utils.utils import ...`` keep working unchanged.
"""
import importlib.util as _ilu
import os as _os
import sys as _sys

import torch

_pkg_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _pkg_root not in _sys.path:
    _sys.path.insert(0, _pkg_root)

from b200cv import yolo_ops as _yolo_ops  # noqa: E402
from b200cv.lib import require_cuda as _require_cuda  # noqa: E402

_ref = _os.environ.get("B200CV_REFERENCE_ROOT")
if _ref:
    _f = _os.path.join(_ref, "CVC-YOLOv3", "utils", "utils.py")
    if _os.path.isfile(_f):
        _spec = _ilu.spec_from_file_location("_b200cv_reference_utils", _f)
        _mod = _ilu.module_from_spec(_spec)
        _spec.loader.exec_module(_mod)
        for _k, _v in vars(_mod).items():
            if not _k.startswith("_") and _k not in ("build_targets", "bbox_iou", "weights_init_normal"):
                globals()[_k] = _v


def weights_init_normal(m):
    name = type(m).__name__
    if "Conv" in name:
        torch.nn.init.normal_(m.weight.data, 0.0, 0.02)
    elif "BatchNorm2d" in name:
        torch.nn.init.normal_(m.weight.data, 1.0, 0.02)
        torch.nn.init.constant_(m.bias.data, 0.0)


def bbox_iou(box1, box2, x1y1x2y2=True):
    """IoU with the reference's '+1 pixel' convention; boxes are (..., 4)."""
    if x1y1x2y2:
        ax1, ay1, ax2, ay2 = box1[..., 0], box1[..., 1], box1[..., 2], box1[..., 3]
        bx1, by1, bx2, by2 = box2[..., 0], box2[..., 1], box2[..., 2], box2[..., 3]
    else:
        ax1, ax2 = box1[..., 0] - box1[..., 2] / 2, box1[..., 0] + box1[..., 2] / 2
        ay1, ay2 = box1[..., 1] - box1[..., 3] / 2, box1[..., 1] + box1[..., 3] / 2
        bx1, bx2 = box2[..., 0] - box2[..., 2] / 2, box2[..., 0] + box2[..., 2] / 2
        by1, by2 = box2[..., 1] - box2[..., 3] / 2, box2[..., 1] + box2[..., 3] / 2
    iw = torch.clamp(torch.min(ax2, bx2) - torch.max(ax1, bx1) + 1, min=0)
    ih = torch.clamp(torch.min(ay2, by2) - torch.max(ay1, by1) + 1, min=0)
    inter = iw * ih
    area_a = (ax2 - ax1 + 1) * (ay2 - ay1 + 1)
    area_b = (bx2 - bx1 + 1) * (by2 - by1 + 1)
    return inter / (area_a + area_b - inter + 1e-12)


def build_targets(target, anchors, num_anchors, num_classes, grid_size_h, grid_size_w, ignore_thres):
    """Same eight tensors as the reference: mask, conf_mask (uint8), tx, ty, tw, th, tconf (float32),
    tcls (uint8) -- computed on the GPU (duplicate cells: the last (b, t) in row-major order wins)."""
    _require_cuda(target, "build_targets")
    anchors = anchors.to(device=target.device, dtype=torch.float32).contiguous()
    if anchors.shape[0] != num_anchors:
        raise ValueError("num_anchors does not match anchors")
    yt = _yolo_ops.yolo_targets(target, anchors, grid_size_h, grid_size_w, ignore_thres)
    return _yolo_ops.yolo_targets_dense(yt, num_classes)
