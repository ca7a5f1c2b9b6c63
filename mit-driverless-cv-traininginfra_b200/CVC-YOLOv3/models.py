"""B200-native drop-in for the reference's CVC-YOLOv3/models.py.

Same public surface -- ``Darknet(config_path, xy_loss, wh_loss, no_object_loss, object_loss, vanilla_anchor)``,
``YOLOLayer(...)``, ``create_modules``, ``vanilla_anchor_list``, the getters, ``load_weights`` /
``save_weights`` and the ``module_list.<i>.conv_<i>`` / ``batch_norm_<i>`` parameter names -- but
``forward`` never touches ATen/cuDNN: it hands the whole network to ``b200cv.darknet_engine`` (tcgen05
implicit-GEMM convolutions, fused BN/LeakyReLU passes, one-kernel YOLO loss).  Tensors must live on a
CUDA (sm_100a) device; there is no CPU path.

Reference behaviour each piece mirrors is cited as models.py:<line> of the reference.
"""
import csv
import os
import sys
from datetime import datetime  # noqa: F401  (kept: scripts reach it through this module)

import numpy as np
import torch
import torch.nn as nn

_pkg_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _pkg_root not in sys.path:
    sys.path.insert(0, _pkg_root)

from utils.parse_config import parse_model_config  # noqa: E402
from utils.utils import build_targets  # noqa: E402,F401

from b200cv import parallel as _parallel  # noqa: E402

_parallel.bind_process_to_local_gpu()  # B200CV_AUTO_DP=1 under torchrun: one GPU per process, before CUDA starts

from b200cv import yolo_ops  # noqa: E402
from b200cv.darknet_engine import DarknetEngine  # noqa: E402

_DP_MESSAGE = (
    "b200cv: nn.DataParallel (one process driving several GPUs) is not supported by the B200 engine.  The reference "
    "wraps the model whenever torch.cuda.device_count() > 1 (train.py:193-195); run one process per GPU instead: "
    "`B200CV_AUTO_DP=1 torchrun --nproc-per-node <N> train.py ...` keeps the script unchanged (every process binds to "
    "its own GPU, so device_count() == 1 and nothing is wrapped; Darknet.forward takes that rank's DataParallel shard "
    "of each batch and the gradients are SUMmed over NCCL like DataParallel's reduce-add), or expose a single GPU with "
    "CUDA_VISIBLE_DEVICES.  See INTEGRATION.md.")

vanilla_anchor_list = [[10, 13], [16, 30], [33, 23], [30, 61], [62, 45], [59, 119], [116, 90], [156, 198], [373, 326]]


def _anchors_from_csv(csv_uri):
    # row 0 of train.csv is a single cell "w,h|w,h|..." (models.py:29-35)
    with open(csv_uri) as f:
        first = next(csv.reader(f))
    cell = str(first)[2:-2]
    return [[float(v) for v in pair.split(",")] for pair in cell.split("'")[0].split("|")]


class EmptyLayer(nn.Module):
    """Placeholder for 'route' and 'shortcut' blocks."""


def create_modules(module_defs, xy_loss, wh_loss, no_object_loss, object_loss, vanilla_anchor):
    """cfg blocks -> (hyperparams, nn.ModuleList) with the reference's module/parameter names
    (models.py:15-110).  Pops the [net] block off ``module_defs`` like the reference does."""
    hyperparams = module_defs.pop(0)
    num_classes = int(hyperparams["classes"])
    img_width, img_height = int(hyperparams["width"]), int(hyperparams["height"])
    int(hyperparams["onnx_height"])  # must exist (models.py:23)
    leaky_slope = float(hyperparams["leaky_slope"])
    activation = hyperparams["conv_activation"]
    masks = [[int(v) for v in group.split(",")] for group in hyperparams["yolo_masks"].split("|")]
    anchor_list = _anchors_from_csv(hyperparams["train_uri"])
    if vanilla_anchor:
        anchor_list = vanilla_anchor_list
    ignore_thresh = float(hyperparams["build_targets_ignore_thresh"])

    widths = [int(hyperparams["channels"])]  # widths[k+1] = channels produced by block k
    module_list = nn.ModuleList()
    head = 0
    linear_next = False
    filters = widths[0]
    for i, d in enumerate(module_defs):
        seq = nn.Sequential()
        kind = d["type"]
        if kind == "convolutional":
            is_head = d["filters"] == "preyolo"
            filters = (num_classes + 5) * len(masks[head]) if is_head else int(d["filters"])
            k = int(d["size"])
            seq.add_module("conv_%d" % i, nn.Conv2d(widths[-1], filters, k, int(d["stride"]), (k - 1) // 2,
                                                     bias=is_head))
            if not is_head:
                seq.add_module("batch_norm_%d" % i, nn.BatchNorm2d(filters))
                if activation == "leaky":
                    seq.add_module("leaky_%d" % i, nn.LeakyReLU(leaky_slope))
                elif activation == "ReLU":
                    seq.add_module("ReLU_%d" % i, nn.ReLU())
        elif kind == "maxpool":
            k, s = int(d["size"]), int(d["stride"])
            if k == 2 and s == 1:
                seq.add_module("_debug_padding_%d" % i, nn.ZeroPad2d((0, 1, 0, 1)))
            seq.add_module("maxpool_%d" % i, nn.MaxPool2d(k, s, (k - 1) // 2))
        elif kind == "upsample":
            seq.add_module("upsample_%d" % i, nn.Upsample(scale_factor=int(d["stride"]), mode="nearest"))
        elif kind == "route":
            filters = 0
            for j in (int(v) for v in d["layers"].split(",")):
                filters += widths[j + 1 if j > 0 else j]
            seq.add_module("route_%d" % i, EmptyLayer())
        elif kind == "shortcut":
            filters = widths[int(d["from"])]
            seq.add_module("shortcut_%d" % i, EmptyLayer())
        elif kind == "yolo":
            anchors = [anchor_list[j] for j in masks[head]]
            seq.add_module("yolo_%d" % i, YOLOLayer(anchors, num_classes, img_height, img_width, ignore_thresh,
                                                    activation, xy_loss, wh_loss, object_loss, no_object_loss))
            head += 1
        module_list.append(seq)
        widths.append(filters)
    return hyperparams, module_list


class YOLOLayer(nn.Module):
    """Detection layer (models.py:118-220): grid/anchor decode, target assignment and the multi-part
    loss, each a single CUDA kernel.  ``forward(sample, targets)`` returns ``(loss, 6-vector)`` with
    the parts ordered (x, y, w, h, obj, noobj); ``forward(sample)`` returns [B, A*G*G, 5+C] detections."""

    def __init__(self, anchors, num_classes, img_height, img_width, build_targets_ignore_thresh, conv_activation,
                 xy_loss, wh_loss, object_loss, no_object_loss):
        super().__init__()
        self.anchors = anchors
        self.num_anchors = len(anchors)
        self.num_classes = num_classes
        self.bbox_attrs = 5 + num_classes
        self.image_height = img_height
        self.image_width = img_width
        self.ignore_thres = build_targets_ignore_thresh
        self.xy_loss = xy_loss
        self.wh_loss = wh_loss
        self.no_object_loss = no_object_loss
        self.object_loss = object_loss
        self.conv_activation = conv_activation

    def forward(self, sample, targets=None):
        n_gh = sample.size(2)
        stride = self.image_height / n_gh  # from the cfg, not the tensor (models.py:145)
        sa = yolo_ops.scaled_anchors(self.anchors, stride, sample.device)
        if targets is not None:
            consts = (self.xy_loss, self.wh_loss, self.object_loss, self.no_object_loss)
            out7 = yolo_ops.YoloLayerFn.apply(sample, targets, sa, self.num_classes, self.ignore_thres, consts)
            return out7[0], out7[1:].detach()
        z = sample.detach().contiguous().float()
        rows = self.num_anchors * n_gh * sample.size(3)
        out = torch.empty(z.shape[0], rows, self.bbox_attrs, dtype=torch.float32, device=z.device)
        yolo_ops.yolo_decode(z, True, self.num_anchors, self.num_classes, sa, stride, out, 0)
        return out


class Darknet(nn.Module):
    """YOLOv3 object detection model (models.py:222-422)."""

    def __init__(self, config_path, xy_loss, wh_loss, no_object_loss, object_loss, vanilla_anchor):
        super().__init__()
        self.module_defs = parse_model_config(config_path)
        self.hyperparams, self.module_list = create_modules(
            module_defs=self.module_defs, xy_loss=xy_loss, wh_loss=wh_loss, no_object_loss=no_object_loss,
            object_loss=object_loss, vanilla_anchor=vanilla_anchor)
        hp = self.hyperparams
        self.img_width, self.img_height = int(hp["width"]), int(hp["height"])
        self.onnx_height = int(hp["onnx_height"])
        self.onnx_name = "%s_%d%d.onnx" % (config_path.split("/")[-1].split(".")[0], self.img_width, self.onnx_height)
        self.num_classes = int(hp["classes"])
        channels = int(hp["channels"])
        if channels not in (1, 3):
            print("Channels in cfg file is not set properly, making it colour")
        self.bw = channels == 1
        self.validate_uri, self.train_uri = hp["validate_uri"], hp["train_uri"]
        self.num_train_images, self.num_validate_images = int(hp["num_train_images"]), int(hp["num_validate_images"])
        self.conf_thresh, self.nms_thresh, self.iou_thresh = (float(hp["conf_thresh"]), float(hp["nms_thresh"]),
                                                              float(hp["iou_thresh"]))
        self.start_weights_dim = [int(v) for v in hp["start_weights_dim"].split(",")]
        self.conv_activation = hp["conv_activation"]
        self.xy_loss, self.wh_loss = xy_loss, wh_loss
        self.no_object_loss, self.object_loss = no_object_loss, object_loss
        self.anchors = vanilla_anchor_list if vanilla_anchor else _anchors_from_csv(self.train_uri)
        self.seen = 0
        self.header_info = torch.tensor([0, 0, 0, self.seen, 0])
        self._engine = None

    # -- getters used by train.py:101-120 -------------------------------------------------------
    def get_start_weight_dim(self):
        return self.start_weights_dim

    def get_onnx_name(self):
        return self.onnx_name

    def get_bw(self):
        return self.bw

    def get_loss_constant(self):
        return [self.xy_loss, self.wh_loss, self.no_object_loss, self.object_loss]

    def get_conv_activation(self):
        return self.conv_activation

    def get_num_classes(self):
        return self.num_classes

    def get_anchors(self):
        return self.anchors

    def get_threshs(self):
        return self.conf_thresh, self.nms_thresh, self.iou_thresh

    def img_size(self):
        return self.img_width, self.img_height

    def get_links(self):
        return self.validate_uri, self.train_uri

    def num_images(self):
        return self.num_validate_images, self.num_train_images

    # -- the hot path --------------------------------------------------------------------------
    def engine(self):
        if self._engine is None:
            object.__setattr__(self, "_engine", DarknetEngine(self))
        return self._engine

    def _replicate_for_data_parallel(self):
        raise RuntimeError(_DP_MESSAGE)

    def forward(self, x, targets=None):
        """Training (targets given): 7-tuple of 0-dim tensors (total, x, y, w, h, obj, noobj), the
        first differentiable (models.py:338).  Inference: [B, sum A*G*G, 5+C] detections."""
        eng = self.engine()
        if x.is_cuda and next(self.parameters()).device != x.device:
            raise RuntimeError(f"Darknet.forward: input on {x.device} but the parameters are on "
                               f"{next(self.parameters()).device}. " + _DP_MESSAGE)
        if targets is not None:
            if _parallel.auto_dp_enabled():
                # DataParallel-equivalent step under torchrun: this rank's scatter chunk of the batch, per-replica
                # mean loss, the returned values summed over the replicas (train.py:70 `losses[0].sum()`)
                _parallel.ensure_group()
                lo, hi = _parallel.dp_chunk(x.shape[0])
                if hi <= lo:
                    raise RuntimeError("B200CV_AUTO_DP: batch smaller than the number of processes")
                out7 = _parallel.sum_over_replicas(eng.train_forward(x[lo:hi].contiguous(), targets[lo:hi].contiguous()))
            else:
                out7 = eng.train_forward(x, targets)
            parts = out7.detach()
            return (out7[0], parts[1], parts[2], parts[3], parts[4], parts[5], parts[6])
        return eng.detect(x)

    # -- Darknet .weights I/O (models.py:339-422): 5 x int32 header, then per conv
    #    [bn.bias, bn.weight, running_mean, running_var] or [conv.bias], then conv.weight (OIHW fp32)
    def load_weights(self, weights_path, start_weight_dim):
        with open(weights_path, "rb") as fp:
            header = np.fromfile(fp, dtype=np.int32, count=5)
            blob = np.fromfile(fp, dtype=np.float32)
        self.header_info = header
        self.seen = header[3]
        pos = 0

        def take(dst, n_file=None):
            nonlocal pos
            n = dst.numel()
            dst.data.copy_(torch.from_numpy(blob[pos:pos + n]).view_as(dst))
            pos += n if n_file is None else n_file

        head = 0
        for d, m in zip(self.module_defs, self.module_list):
            if d["type"] != "convolutional":
                continue
            conv = m[0]
            if d["filters"] != "preyolo":
                bn = m[1]
                take(bn.bias)
                take(bn.weight)
                take(bn.running_mean)
                take(bn.running_var)
                take(conv.weight)
            else:
                # the file holds `orig` output filters (e.g. 255); keep the first num_b of them
                orig = start_weight_dim[head]
                head += 1
                num_b = conv.bias.numel()
                take(conv.bias, n_file=orig)
                per_filter = conv.weight.numel() // num_b
                w = torch.from_numpy(blob[pos:pos + per_filter * orig]).view(orig, *conv.weight.shape[1:])
                conv.weight.data.copy_(w[:num_b])
                pos += per_filter * orig

    def save_weights(self, path, cutoff=-1):
        with open(path, "wb") as fp:
            self.header_info[3] = self.seen
            header = self.header_info
            if isinstance(header, torch.Tensor):  # the reference only works after a load (numpy header)
                header = header.cpu().numpy().astype(np.int32)
            header.tofile(fp)
            for d, m in zip(self.module_defs[:cutoff], self.module_list[:cutoff]):
                if d["type"] != "convolutional":
                    continue
                conv = m[0]
                if d["filters"] != "preyolo":
                    bn = m[1]
                    for t in (bn.bias, bn.weight, bn.running_mean, bn.running_var):
                        t.data.cpu().numpy().tofile(fp)
                else:
                    conv.bias.data.cpu().numpy().tofile(fp)
                conv.weight.data.cpu().numpy().tofile(fp)
