"""Programmatic Darknet cfg writers for the two networks the reference ships
(CVC-YOLOv3/model_cfg/yolo_baseline.cfg = Darknet-53 + 3 heads, yolo_baseline_tiny.cfg = 2 heads).

The reference's cfg files cannot travel to the GPU box, so tests and bench.py generate the same
architectures from code (tests/test_host_logic.py checks block-for-block equality with the shipped
files when the reference tree is available).  The [net] keys are the ones models.Darknet reads.
"""
from __future__ import annotations

import os

VANILLA_ANCHORS = [[10, 13], [16, 30], [33, 23], [30, 61], [62, 45], [59, 119], [116, 90], [156, 198], [373, 326]]


def _net(width, height, classes, masks, scales, train_uri, start_dims):
    return "\n".join([
        "[net]", f"width={width}", f"height={height}", "onnx_height=320", f"classes={classes}", "channels=3",
        f"yolo_masks={masks}", f"yolo_scales={scales}", "validate_uri=dataset/validate.csv", f"train_uri={train_uri}",
        "weights_uri=none.weights", f"start_weights_dim={start_dims}", "num_train_images=-1",
        "num_validate_images=-1", "leaky_slope=0.1", "conv_activation=leaky", "build_targets_ignore_thresh=0.5",
        "conf_thresh=0.8", "nms_thresh=0.25", "iou_thresh=0.5", ""])


def _conv(filters, size, stride=1):
    return f"[convolutional]\nfilters={filters}\nsize={size}\nstride={stride}\n"


def _head():
    return "[convolutional]\nsize=1\nstride=1\nfilters=preyolo\nactivation=linear\n\n[yolo]\nnote=head\n"


def darknet53_cfg(width=416, height=416, classes=80, train_uri="train.csv") -> str:
    out = [_net(width, height, classes, "6,7,8|3,4,5|0,1,2", "32,16,8", train_uri, "255,255,255")]
    out.append(_conv(32, 3, 1))
    for ch, reps in ((64, 1), (128, 2), (256, 8), (512, 8), (1024, 4)):
        out.append(_conv(ch, 3, 2))  # downsample
        for _ in range(reps):
            out += [_conv(ch // 2, 1), _conv(ch, 3), "[shortcut]\nfrom=-3\nactivation=linear\n"]
    for ch, route_to in ((512, None), (256, 61), (128, 36)):
        if route_to is not None:
            out += ["[route]\nlayers = -4\n", _conv(ch, 1), "[upsample]\nstride=2\n", f"[route]\nlayers = -1, {route_to}\n"]
        for _ in range(3):
            out += [_conv(ch, 1), _conv(ch * 2, 3)]
        out.append(_head())
    return "\n".join(out)


def tiny_cfg(width=416, height=416, classes=80, train_uri="train.csv") -> str:
    out = [_net(width, height, classes, "3,4,5|0,1,2", "32,16", train_uri, "255,255")]
    for ch in (16, 32, 64, 128, 256):
        out += [_conv(ch, 3), "[maxpool]\nsize=2\nstride=2\n"]
    out += [_conv(512, 3), "[maxpool]\nsize=2\nstride=1\n", _conv(1024, 3), _conv(256, 1), _conv(512, 3), _head()]
    out += ["[route]\nlayers = -4\n", _conv(128, 1), "[upsample]\nstride=2\n", "[route]\nlayers = -1, 8\n",
            _conv(256, 3), _head()]
    return "\n".join(out)


def write_anchor_csv(path, anchors=VANILLA_ANCHORS):
    """train.csv whose row 0 is the single anchors cell models.Darknet parses (models.py:29-35)."""
    with open(path, "w") as f:
        f.write('"' + "|".join(f"{a},{b}" for a, b in anchors) + '"\n')
        f.write("Name,URL,Width,Height,Scale,X0 Y0 H0 W0\n")


def write_cfg(directory, kind, width, height, classes) -> str:
    """Write <directory>/<kind>_<w>x<h>_c<classes>.cfg (+ train.csv); returns the cfg path."""
    os.makedirs(directory, exist_ok=True)
    csv_path = os.path.join(directory, "train.csv")
    if not os.path.exists(csv_path):
        write_anchor_csv(csv_path)
    text = {"darknet53": darknet53_cfg, "tiny": tiny_cfg}[kind](width, height, classes, csv_path)
    path = os.path.join(directory, f"{kind}_{width}x{height}_c{classes}.cfg")
    with open(path, "w") as f:
        f.write(text)
    return path
