"""Residual block of RektNet (reference RektNet/resnet.py:8-27).

relu( BN(1x1(x)) + BN(3x3( relu(BN(3x3, dilation 2, pad 2 (x))) )) ), all convs with bias, stride 1.
Inside KeypointNet the blocks are executed by b200cv.rektnet_engine.RektNetEngine (NHWC bf16 end to end); a block
called on its own (NCHW fp32 in and out, differentiable w.r.t. its input) runs through
b200cv.rektnet_engine.ResBlockEngine on the same kernels.
"""
import os
import sys

import torch.nn as nn

_pkg_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _pkg_root not in sys.path:
    sys.path.insert(0, _pkg_root)


class ResNet(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv1 = nn.Conv2d(in_channels=in_channels, out_channels=out_channels, kernel_size=3, stride=1, padding=2,
                               dilation=2)
        self.bn1 = nn.BatchNorm2d(out_channels)
        self.relu1 = nn.ReLU()
        self.conv2 = nn.Conv2d(in_channels=out_channels, out_channels=out_channels, kernel_size=3, stride=1, padding=1)
        self.bn2 = nn.BatchNorm2d(out_channels)
        self.relu2 = nn.ReLU()
        self.shortcut_conv = nn.Conv2d(in_channels=in_channels, out_channels=out_channels, kernel_size=1, stride=1)
        self.shortcut_bn = nn.BatchNorm2d(out_channels)

        self._engine = None

    def forward(self, x):
        """relu(bn_s(1x1(x)) + bn_2(3x3(relu(bn_1(3x3 dilated(x)))))) on the B200 kernels (reference :22-27)."""
        if self._engine is None:
            from b200cv.rektnet_engine import ResBlockEngine

            object.__setattr__(self, "_engine", ResBlockEngine(self))
        return self._engine.run(x)
