"""ORACLE (test infrastructure, never shipped, never on the product path).

CPU fp32 restatement of the CVC-YOLOv3 hot path of cv-core/MIT-Driverless-CV-TrainingInfra, written
as a *functional interpreter* over a dict of named tensors (the reference is an nn.ModuleList
interpreter).  Every function cites the reference lines it follows (paths relative to the
reference tree).  Parity is PINNED: tests/test_oracle_golden.py checks this file against vectors
produced by importing the reference itself (oracle/gen_golden.py -> tests/golden/).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
from __future__ import annotations

import csv
import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

VANILLA_ANCHORS = [[10, 13], [16, 30], [33, 23], [30, 61], [62, 45], [59, 119], [116, 90], [156, 198],
                   [373, 326]]  # CVC-YOLOv3/models.py:13


# --------------------------------------------------------------------------- cfg parsing
def parse_cfg(path: str) -> List[dict]:
    """CVC-YOLOv3/utils/parse_config.py:1-18 -- blocks of key=value; '#' and empty lines dropped."""
    blocks: List[dict] = []
    with open(path) as f:
        for raw in f.read().split("\n"):
            if not raw or raw.startswith("#"):  # checked BEFORE stripping, like the reference (:5)
                continue
            line = raw.strip()
            if line.startswith("["):
                blocks.append({"type": line[1:-1].rstrip()})
                if blocks[-1]["type"] == "convolutional":
                    blocks[-1]["batch_normalize"] = 0  # :12-13
            else:
                key, value = line.split("=")
                blocks[-1][key.rstrip()] = value.strip()
    return blocks


def read_anchor_row(csv_path: str) -> List[List[float]]:
    """CVC-YOLOv3/models.py:29-35 -- row 0 of train.csv is ONE cell 'w,h|w,h|...'."""
    with open(csv_path) as f:
        row = next(csv.reader(f))
    cell = str(row)[2:-2]
    return [[float(y) for y in x.split(",")] for x in cell.split("'")[0].split("|")]


class NetSpec:
    """Static description of a Darknet cfg: per-layer dicts + hyper-parameters (models.py:15-110)."""

    def __init__(self, cfg_path: str, vanilla_anchor: bool = True):
        defs = parse_cfg(cfg_path)
        self.hyper = defs.pop(0)  # :19
        self.defs = defs
        h = self.hyper
        self.width, self.height = int(h["width"]), int(h["height"])
        self.num_classes = int(h["classes"])
        self.channels = int(h["channels"])
        self.leaky_slope = float(h["leaky_slope"])
        self.conv_activation = h["conv_activation"]
        self.ignore_thresh = float(h["build_targets_ignore_thresh"])
        self.yolo_masks = [[int(y) for y in x.split(",")] for x in h["yolo_masks"].split("|")]
        self.anchors = VANILLA_ANCHORS if vanilla_anchor else read_anchor_row(h["train_uri"])
        # channel bookkeeping, :44-109
        out_filters = [self.channels]
        self.layers: List[dict] = []
        yolo_count = 0
        act_flag = True
        for i, d in enumerate(defs):
            L = {"type": d["type"], "index": i}
            if d["type"] == "convolutional":
                if d["filters"] == "preyolo":  # :51-54
                    filters = (self.num_classes + 5) * len(self.yolo_masks[yolo_count])
                    L["bn"], act_flag = False, False
                else:
                    filters = int(d["filters"])
                    L["bn"] = True
                k = int(d["size"])
                L.update(cin=out_filters[-1], cout=filters, k=k, stride=int(d["stride"]), pad=(k - 1) // 2)
                L["act"] = self.conv_activation if act_flag else "linear"  # :68-72
                act_flag = True
            elif d["type"] == "maxpool":
                L.update(k=int(d["size"]), stride=int(d["stride"]))
                filters = out_filters[-1]
            elif d["type"] == "upsample":
                L.update(stride=int(d["stride"]))
                filters = out_filters[-1]
            elif d["type"] == "route":
                idx = [int(x) for x in d["layers"].split(",")]
                L["layers"] = idx
                filters = sum(out_filters[j + 1 if j > 0 else j] for j in idx)  # :93-96
            elif d["type"] == "shortcut":
                L["from"] = int(d["from"])
                filters = out_filters[int(d["from"])]  # :99 (sic: indexes the shifted list)
            elif d["type"] == "yolo":
                L["anchors"] = [self.anchors[j] for j in self.yolo_masks[yolo_count]]
                yolo_count += 1
                # `filters` keeps the previous value, like the reference's loop variable
            L["filters"] = filters
            self.layers.append(L)
            out_filters.append(filters)


def init_params(spec: NetSpec, seed: int = 0) -> Tuple[Dict[str, torch.Tensor], Dict[str, torch.Tensor]]:
    """Parameters/buffers under the reference's state_dict names, initialised like
    weights_init_normal (CVC-YOLOv3/utils/utils.py:50-56) from a seeded CPU generator."""
    g = torch.Generator().manual_seed(seed)
    params: Dict[str, torch.Tensor] = {}
    buffers: Dict[str, torch.Tensor] = {}
    for L in spec.layers:
        if L["type"] != "convolutional":
            continue
        i = L["index"]
        p = f"module_list.{i}."
        params[p + f"conv_{i}.weight"] = torch.empty(L["cout"], L["cin"], L["k"], L["k"]).normal_(0.0, 0.02, generator=g)
        if L["bn"]:
            params[p + f"batch_norm_{i}.weight"] = torch.empty(L["cout"]).normal_(1.0, 0.02, generator=g)
            params[p + f"batch_norm_{i}.bias"] = torch.zeros(L["cout"])
            buffers[p + f"batch_norm_{i}.running_mean"] = torch.zeros(L["cout"])
            buffers[p + f"batch_norm_{i}.running_var"] = torch.ones(L["cout"])
        else:
            bound = 1.0 / math.sqrt(L["cin"] * L["k"] * L["k"])
            params[p + f"conv_{i}.bias"] = torch.empty(L["cout"]).uniform_(-bound, bound, generator=g)
    return params, buffers


# --------------------------------------------------------------------------- target assignment
def bbox_iou_corner(b1: torch.Tensor, b2: torch.Tensor) -> torch.Tensor:
    """CVC-YOLOv3/utils/utils.py:163-193, the x1y1x2y2 branch with its '+1 pixel' convention."""
    ix1, iy1 = torch.max(b1[..., 0], b2[..., 0]), torch.max(b1[..., 1], b2[..., 1])
    ix2, iy2 = torch.min(b1[..., 2], b2[..., 2]), torch.min(b1[..., 3], b2[..., 3])
    inter = torch.clamp(ix2 - ix1 + 1, min=0) * torch.clamp(iy2 - iy1 + 1, min=0)
    a1 = (b1[..., 2] - b1[..., 0] + 1) * (b1[..., 3] - b1[..., 1] + 1)
    a2 = (b2[..., 2] - b2[..., 0] + 1) * (b2[..., 3] - b2[..., 1] + 1)
    return inter / (a1 + a2 - inter + 1e-12)


def build_targets(target: torch.Tensor, anchors: torch.Tensor, num_classes: int, gh: int, gw: int,
                  ignore_thres: float):
    """CVC-YOLOv3/utils/utils.py:195-275.  Returns the same eight tensors, same dtypes.

    Quirks kept on purpose: padding rows become copies of row 0 (:223-228); the ignore mask is
    cleared for ALL images and anchors (:255); duplicate cells -> last (b,t) wins (index_put on CPU)."""
    nb, nt = target.shape[0], target.shape[1]
    na = anchors.shape[0]
    shape = (nb, na, gh, gw)
    mask = torch.zeros(shape, dtype=torch.uint8)
    conf_mask = torch.ones(shape, dtype=torch.uint8)
    tx, ty, tw, th, tconf = (torch.zeros(shape) for _ in range(5))
    tcls = torch.zeros(shape + (num_classes,), dtype=torch.uint8)

    valid = target.sum(dim=2) > 0  # :210
    src = torch.where(valid.unsqueeze(-1), target, target[:, :1, :].expand_as(target))
    gx, gy = src[:, :, 1] * gw, src[:, :, 2] * gh  # :213-216
    gwid, ghei = src[:, :, 3] * gw, src[:, :, 4] * gh
    gi, gj = gx.long(), gy.long()  # :219-220 (truncation)

    zeros = torch.zeros_like(gwid)
    gt_box = torch.stack((zeros, zeros, gwid, ghei), dim=-1).unsqueeze(2)  # [B,T,1,4]
    an_box = torch.cat((torch.zeros(na, 2), anchors), dim=1).view(1, 1, na, 4)
    ious = bbox_iou_corner(gt_box.expand(-1, -1, na, -1), an_box.expand(nb, nt, -1, -1)).permute(0, 2, 1)  # [B,A,T]

    over = ious > ignore_thres  # :244-245
    gj_over = gj.unsqueeze(1).expand(-1, na, -1)[over]
    gi_over = gi.unsqueeze(1).expand(-1, na, -1)[over]
    conf_mask[:, :, gj_over, gi_over] = 0  # :255 -- every image, every anchor
    best = torch.argmax(ious, dim=1)  # :257
    b_idx = torch.arange(nb).view(-1, 1).expand_as(best)

    mask[b_idx, best, gj, gi] = 1
    conf_mask[b_idx, best, gj, gi] = 1
    tx[b_idx, best, gj, gi] = gx - gi.float()
    ty[b_idx, best, gj, gi] = gy - gj.float()
    tw[b_idx, best, gj, gi] = torch.log(gwid / anchors[best, 0] + 1e-16)
    th[b_idx, best, gj, gi] = torch.log(ghei / anchors[best, 1] + 1e-16)
    tcls[b_idx, best, gj, gi, target[:, :, 0].long()] = 1  # :271 label of the row itself
    tconf[b_idx, best, gj, gi] = 1
    return mask, conf_mask, tx, ty, tw, th, tconf, tcls


# --------------------------------------------------------------------------- YOLO layer
def scaled_anchors(anchors: Sequence[Sequence[float]], stride: float) -> torch.Tensor:
    """CVC-YOLOv3/models.py:160 -- divided in Python float64, then cast to fp32."""
    return torch.tensor([(a_w / stride, a_h / stride) for a_w, a_h in anchors], dtype=torch.float)


def yolo_layer(sample: torch.Tensor, targets: Optional[torch.Tensor], anchors, num_classes: int, img_height: int,
               ignore_thres: float, xy_loss: float, wh_loss: float, object_loss: float, no_object_loss: float):
    """CVC-YOLOv3/models.py:140-220.  Training: (loss, 6-vector (x,y,w,h,obj,noobj)); eval: detections."""
    nb, _, ngh, ngw = sample.shape
    na = len(anchors)
    attrs = 5 + num_classes
    stride = img_height / ngh  # :145
    pred = sample.view(nb, na, attrs, ngh, ngw).permute(0, 1, 3, 4, 2).contiguous()  # :147
    x, y = torch.sigmoid(pred[..., 0]), torch.sigmoid(pred[..., 1])
    w, h = pred[..., 2], pred[..., 3]
    conf, cls = torch.sigmoid(pred[..., 4]), torch.sigmoid(pred[..., 5:])
    sa = scaled_anchors(anchors, stride)
    if targets is None:  # :213-220
        gx = torch.arange(ngw, dtype=torch.float).view(1, 1, 1, ngw)
        gy = torch.arange(ngh, dtype=torch.float).view(1, 1, ngh, 1)
        boxes = torch.stack((x.data + gx, y.data + gy, torch.exp(w.data) * sa[:, 0].view(1, na, 1, 1),
                             torch.exp(h.data) * sa[:, 1].view(1, na, 1, 1)), dim=-1)
        return torch.cat((boxes.view(nb, -1, 4) * stride, conf.view(nb, -1, 1), cls.view(nb, -1, num_classes)), -1)
    mask, conf_mask, tx, ty, tw, th, tconf, tcls = build_targets(targets.float(), sa, num_classes, ngh, ngw, ignore_thres)
    # no-op in fp32 (the reference's arithmetic).  Tests also evaluate this restatement in fp64 -- same assignment,
    # wider arithmetic -- as the yardstick for how far the reference's OWN fp32 rounding moves losses and gradients.
    tx, ty, tw, th, tconf = (t.to(sample.dtype) for t in (tx, ty, tw, th, tconf))
    m = mask.bool()
    cf = (conf_mask - mask).bool()  # :196
    mse = lambda a, b: F.mse_loss(a, b, reduction="mean")
    bce = lambda a, b: F.binary_cross_entropy(a, b, reduction="mean")  # on probabilities, :137,207-208
    lx, ly = xy_loss * mse(x[m], tx[m]), xy_loss * mse(y[m], ty[m])
    lw, lh = wh_loss * mse(w[m], tw[m]), wh_loss * mse(h[m], th[m])
    lcls = 0 * (1 / nb) * F.cross_entropy(cls[m], torch.argmax(tcls[m], 1))  # :204-205, weight 0
    lnoobj = no_object_loss * bce(conf[cf], tconf[cf])
    lobj = object_loss * bce(conf[m], tconf[m])
    loss = lx + ly + lw + lh + lnoobj + lobj + lcls
    return loss, torch.tensor((lx, ly, lw, lh, lobj, lnoobj))  # :211


# --------------------------------------------------------------------------- Darknet interpreter
class _StoreBf16(torch.autograd.Function):
    """Straight-through bf16 storage rounding (only used by the `emulate_bf16` diagnostic mode)."""

    @staticmethod
    def forward(ctx, t):
        return t.to(torch.bfloat16).float()

    @staticmethod
    def backward(ctx, g):
        return g


def darknet_forward(spec: NetSpec, params: Dict[str, torch.Tensor], buffers: Dict[str, torch.Tensor],
                    x: torch.Tensor, targets: Optional[torch.Tensor], loss_consts=(2.0, 1.6, 25.0, 0.1),
                    training: bool = True, emulate_bf16: bool = False):
    """CVC-YOLOv3/models.py:312-338 with the module bodies of :59-101 inlined.
    loss_consts = (xy, wh, no_object, object) -- the constructor order of Darknet (:225).

    emulate_bf16=True is NOT the reference algorithm: it additionally rounds the tensors the B200 path
    STORES in bf16 (input image, conv weights, BN'd conv outputs, activations) so that tests can separate
    kernel errors from the expected effect of bf16 storage.  Parity claims use emulate_bf16=False."""
    xy, wh, noobj, obj = loss_consts
    q = _StoreBf16.apply if emulate_bf16 else (lambda t: t)
    x = q(x)
    outs: List[torch.Tensor] = []
    yolo_out = []
    totals = torch.zeros(6)
    for L in spec.layers:
        i, t = L["index"], L["type"]
        p = f"module_list.{i}."
        if t == "convolutional":
            x = F.conv2d(x, q(params[p + f"conv_{i}.weight"]), params.get(p + f"conv_{i}.bias"), stride=L["stride"],
                         padding=L["pad"])
            if L["bn"]:
                x = q(x)
                x = F.batch_norm(x, buffers[p + f"batch_norm_{i}.running_mean"],
                                 buffers[p + f"batch_norm_{i}.running_var"], params[p + f"batch_norm_{i}.weight"],
                                 params[p + f"batch_norm_{i}.bias"], training=training, momentum=0.1, eps=1e-5)
            if L["act"] == "leaky":
                x = F.leaky_relu(x, spec.leaky_slope)
            elif L["act"] == "ReLU":
                x = F.relu(x)
            if L["bn"]:
                x = q(x)
        elif t == "maxpool":
            if L["k"] == 2 and L["stride"] == 1:
                x = F.pad(x, (0, 1, 0, 1))  # ZERO padding, :77-79
            x = F.max_pool2d(x, L["k"], L["stride"], (L["k"] - 1) // 2)
        elif t == "upsample":
            x = F.interpolate(x, scale_factor=L["stride"], mode="nearest")
        elif t == "route":
            x = torch.cat([outs[j] for j in L["layers"]], 1)  # :323-324
        elif t == "shortcut":
            x = q(outs[-1] + outs[L["from"]])  # :326-327
        elif t == "yolo":
            if targets is not None:
                x, parts = yolo_layer(x, targets, L["anchors"], spec.num_classes, spec.height, spec.ignore_thresh, xy,
                                      wh, obj, noobj)
                totals = totals + parts
            else:
                x = yolo_layer(x, None, L["anchors"], spec.num_classes, spec.height, spec.ignore_thresh, xy, wh, obj,
                               noobj)
            yolo_out.append(x)
        outs.append(x)
    if targets is not None:
        return (sum(yolo_out), *totals)
    return torch.cat(yolo_out, 1)


# --------------------------------------------------------------------------- synthetic inputs
def synth_targets(B: int, T: int = 16, seed: int = 1) -> torch.Tensor:
    """Cone-like boxes: n in [1,T] per image, (cls=0, cx, cy, w, h) normalised, remaining rows zero."""
    g = torch.Generator().manual_seed(seed)
    t = torch.zeros(B, T, 5)
    for b in range(B):
        n = int(torch.randint(1, T + 1, (1,), generator=g))
        t[b, :n, 1:3] = 0.05 + 0.9 * torch.rand(n, 2, generator=g)
        t[b, :n, 3] = 0.01 + 0.08 * torch.rand(n, generator=g)
        t[b, :n, 4] = 0.02 + 0.16 * torch.rand(n, generator=g)
    return t


def synth_images(B: int, H: int, W: int, seed: int = 0) -> torch.Tensor:
    return torch.rand(B, 3, H, W, generator=torch.Generator().manual_seed(seed))


def write_cfg_copy(src_cfg: str, dst_cfg: str, width: int, height: int, classes: Optional[int], train_csv: str):
    """Copy of a shipped cfg with width/height/classes/train_uri edited (how 416/608 configs are made)."""
    out = []
    for line in open(src_cfg).read().split("\n"):
        key = line.split("=")[0].strip() if "=" in line else None
        if key == "width":
            line = f"width={width}"
        elif key == "height":
            line = f"height={height}"
        elif key == "classes" and classes is not None:
            line = f"classes={classes}"
        elif key == "train_uri":
            line = f"train_uri={train_csv}"
        out.append(line)
    with open(dst_cfg, "w") as f:
        f.write("\n".join(out))


def write_anchor_csv(path: str, anchors=VANILLA_ANCHORS):
    with open(path, "w") as f:
        f.write('"' + "|".join(f"{a},{b}" for a, b in anchors) + '"\n')
        f.write("Name,URL,Width,Height,Scale,X0 Y0 H0 W0\n")
