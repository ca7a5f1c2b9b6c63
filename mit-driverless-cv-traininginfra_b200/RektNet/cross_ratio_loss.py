"""B200-native drop-in for the reference's RektNet/cross_ratio_loss.py.

``CrossRatioLoss(loss_type, include_geo, geo_loss_gamma_horz, geo_loss_gamma_vert)
.forward(heatmap, points, target_hm, target_points)`` -> ``(location_loss, geo_loss, total)``
(reference :20-63).  The forward is two small CUDA kernels; the backward is fused into KeypointNet's
head backward when ``heatmap``/``points`` come straight from a B200 KeypointNet.
"""
import os
import sys

import torch
from torch import nn

_pkg_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _pkg_root not in sys.path:
    sys.path.insert(0, _pkg_root)

from b200cv.rektnet_engine import LOSS_TYPES, CrossRatioLossFn  # noqa: E402


class CrossRatioLoss(nn.Module):
    def __init__(self, loss_type, include_geo, geo_loss_gamma_horz, geo_loss_gamma_vert):
        super().__init__()
        self.loss_type = loss_type
        self.include_geo = include_geo
        self.geo_loss_gamma_vert = geo_loss_gamma_vert
        self.geo_loss_gamma_horz = geo_loss_gamma_horz
        print(f"Including geometric loss: {include_geo}")
        print(f"Loss type: {loss_type}")

    def forward(self, heatmap, points, target_hm, target_points):
        if self.loss_type not in LOSS_TYPES:
            print("Did not recognize loss function selection!")
            sys.exit(1)
        handle = getattr(points, "_b200cv_head", None)
        if handle is not None and getattr(heatmap, "_b200cv_head", None) is not handle:
            handle = None
        loss3 = CrossRatioLossFn.apply(heatmap, points, target_hm, target_points, LOSS_TYPES[self.loss_type],
                                       bool(self.include_geo), float(self.geo_loss_gamma_horz),
                                       float(self.geo_loss_gamma_vert), handle)
        location_loss, total = loss3[0], loss3[2]
        # the reference returns an int64 CPU zero when the geometric term is off (:59)
        geo_loss = loss3[1] if self.include_geo else torch.tensor(0)
        return location_loss, geo_loss, total
