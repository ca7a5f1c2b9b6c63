"""Residual block of RektNet (reference RektNet/resnet.py:8-27): parameter container only.

relu( BN(1x1(x)) + BN(3x3( relu(BN(3x3, dilation 2, pad 2 (x))) )) ), all convs with bias, stride 1.
The arithmetic is executed by b200cv.rektnet_engine on the B200 kernels; calling a block on its own
runs it through a one-block engine.
"""
import torch.nn as nn


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

    def forward(self, x):
        raise RuntimeError("ResNet blocks are executed by KeypointNet's B200 engine; call KeypointNet.forward")
