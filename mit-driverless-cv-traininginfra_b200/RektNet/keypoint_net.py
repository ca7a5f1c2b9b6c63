"""B200-native drop-in for the reference's RektNet/keypoint_net.py.

``KeypointNet(num_kpt=7, image_size=(80, 80), onnx_mode=False, init_weight=True).forward(x)`` returns
``(heatmap [B,K,H,W] softmaxed over H*W, points [B,K,2])`` (or raw logits in onnx_mode) exactly like
the reference (:58-70), with the same state_dict names, but runs on hand-written sm_100a kernels via
b200cv.rektnet_engine.  CUDA tensors only; there is no CPU path.
"""
import os
import sys

import torch  # noqa: F401
import torch.nn as nn

_here = os.path.dirname(os.path.abspath(__file__))
_pkg_root = os.path.dirname(_here)
for _p in (_here, _pkg_root):
    if _p not in sys.path:
        sys.path.insert(0, _p)

from resnet import ResNet  # noqa: E402
from cross_ratio_loss import CrossRatioLoss  # noqa: E402,F401

from b200cv.rektnet_engine import RektNetEngine  # noqa: E402


class KeypointNet(nn.Module):
    def __init__(self, num_kpt=7, image_size=(80, 80), onnx_mode=False, init_weight=True):
        super().__init__()
        net_size = 16
        self.conv = nn.Conv2d(in_channels=3, out_channels=net_size, kernel_size=7, stride=1, padding=3)
        self.bn = nn.BatchNorm2d(net_size)
        self.relu = nn.ReLU()
        self.res1 = ResNet(net_size, net_size)
        self.res2 = ResNet(net_size, net_size * 2)
        self.res3 = ResNet(net_size * 2, net_size * 4)
        self.res4 = ResNet(net_size * 4, net_size * 8)
        self.out = nn.Conv2d(in_channels=net_size * 8, out_channels=num_kpt, kernel_size=1, stride=1, padding=0)
        if init_weight:
            self._initialize_weights()
        self.image_size = image_size
        self.num_kpt = num_kpt
        self.onnx_mode = onnx_mode
        self._engine = None

    def _initialize_weights(self):
        # Kaiming-normal (fan_out) convs, zero biases, BN gamma=1 beta=0 (reference :33-44)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, (nn.BatchNorm2d, nn.GroupNorm)):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, 0, 0.01)
                nn.init.constant_(m.bias, 0)

    def engine(self):
        if self._engine is None:
            object.__setattr__(self, "_engine", RektNetEngine(self))
        return self._engine

    def forward(self, x):
        out = self.engine().run(x)
        if self.onnx_mode:
            return out
        hm, pts = out
        return hm, pts.view(-1, self.num_kpt, 2) if pts.shape[1] != self.num_kpt else pts
