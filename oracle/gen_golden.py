"""Generates tests/golden/*.pt by IMPORTING THE REFERENCE from /root/reference (read-only) and running
it on CPU fp32.  Run in the build container only (the GPU box has no /root/reference); the small
fixtures it writes are committed, and so is this script.

    python oracle/gen_golden.py

Inputs are described by seeds/recipes that tests re-create through oracle/*.py helpers; only
reference OUTPUTS (and small adversarial inputs) are stored.
"""
import importlib
import os
import sys
import tempfile
import warnings

import torch

warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("B200CV_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import rektnet_oracle as RO  # noqa: E402
from oracle import yolo_oracle as YO  # noqa: E402


def import_reference(subdir, names):
    """Import reference modules with <REF>/<subdir> first on sys.path, then restore sys.path/modules."""
    path = os.path.join(REF, subdir)
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "utils" or k.startswith("utils.")}
    sys.path.insert(0, path)
    try:
        mods = [importlib.import_module(n) for n in names]
    finally:
        sys.path.remove(path)
    for k in list(sys.modules):
        if k == "utils" or k.startswith("utils.") or k in names:
            sys.modules.pop(k)
    sys.modules.update(saved)
    return mods


def param_digest(named):
    return {k: (float(v.double().sum()), float(v.double().abs().sum())) for k, v in named}


def sample_grads(named, n=16):
    out = {}
    for k, p in named:
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        flat = g.flatten()
        idx = torch.linspace(0, flat.numel() - 1, min(n, flat.numel())).long()
        out[k] = {"norm": float(g.double().norm()), "idx": idx, "val": flat[idx].clone()}
    return out


# ------------------------------------------------------------------------------------------ YOLO
def adversarial_targets():
    """Small hand-made target sets that hit every build_targets quirk."""
    cases = {}
    t = torch.zeros(2, 4, 5)
    t[0, 0] = torch.tensor([0, 0.31, 0.52, 0.10, 0.20])
    t[0, 1] = torch.tensor([0, 0.33, 0.53, 0.12, 0.22])  # same cell as row 0 at G=13 -> collision
    t[1, 0] = torch.tensor([0, 0.70, 0.20, 0.05, 0.09])
    cases["collision_with_padding"] = t
    t = torch.zeros(2, 2, 5)
    t[0, 0] = torch.tensor([0, 0.31, 0.52, 0.10, 0.20])
    t[0, 1] = torch.tensor([0, 0.33, 0.53, 0.30, 0.40])  # full rows: last writer wins
    t[1, 0] = torch.tensor([0, 0.10, 0.10, 0.05, 0.09])
    t[1, 1] = torch.tensor([0, 0.90, 0.90, 0.05, 0.09])
    cases["collision_no_padding"] = t
    t = torch.zeros(3, 3, 5)
    t[0, 0] = torch.tensor([0, 0.5, 0.5, 0.2, 0.3])
    t[2, 0] = torch.tensor([0, 0.25, 0.75, 0.9, 0.9])  # image 1 has NO target -> phantom at (0,0)
    cases["zero_target_image"] = t
    t = torch.zeros(2, 3, 5)
    t[0, 0] = torch.tensor([0, 0.5, 0.5, 0.6, 0.7])  # large: several anchors over the ignore threshold
    t[0, 1] = torch.tensor([0, 6.0 / 13, 7.0 / 13, 0.1, 0.1])  # exactly on a cell boundary
    t[1, 0] = torch.tensor([0, 0.999, 0.999, 0.02, 0.02])
    t[1, 1] = torch.tensor([2, 0.40, 0.60, 0.3, 0.2])  # non-zero class label
    cases["ignore_boundary_label"] = t
    return cases


def gen_yolo(out):
    models, uu = import_reference("CVC-YOLOv3", ["models", "utils.utils"])
    tmp = tempfile.mkdtemp()
    csv_path = os.path.join(tmp, "train.csv")
    YO.write_anchor_csv(csv_path)

    # (1) build_targets
    bt = {}
    for G in (13, 26):
        stride = 416 / G
        for a_name, mask in (("large", [6, 7, 8]), ("small", [0, 1, 2])):
            anchors = YO.scaled_anchors([YO.VANILLA_ANCHORS[i] for i in mask], stride)
            for name, tg in {**adversarial_targets(), "synth": YO.synth_targets(4, 8, seed=3)}.items():
                for C in (1, 3):
                    if (name == "ignore_boundary_label") != (C == 3):
                        continue  # the label-2 case needs C=3 (the reference raises IndexError otherwise)
                    res = uu.build_targets(tg.clone(), anchors, 3, C, G, G, 0.5)
                    bt[f"{name}_G{G}_{a_name}_C{C}"] = {"targets": tg, "anchors": anchors, "G": G, "C": C,
                                                         "out": [r.clone() for r in res]}
    out["build_targets"] = bt

    # (2) YOLOLayer loss + gradient w.r.t. the raw head
    yl = {}
    for name, (B, C, G, scale, tseed) in {"c1_g13": (4, 1, 13, 1.0, 5), "c1_g26": (2, 1, 26, 1.0, 6),
                                          "c3_g13": (2, 3, 13, 1.0, 7), "c80_g13": (2, 80, 13, 1.0, 8),
                                          "c1_g13_saturated": (2, 1, 13, 30.0, 9)}.items():
        anchors = [YO.VANILLA_ANCHORS[i] for i in (6, 7, 8)]
        layer = models.YOLOLayer(anchors, C, 416, 416, 0.5, "leaky", 2.0, 1.6, 0.1, 25.0)
        g = torch.Generator().manual_seed(100 + tseed)
        sample = (torch.randn(B, 3 * (5 + C), G, G, generator=g) * scale).requires_grad_(True)
        tg = YO.synth_targets(B, 8, seed=tseed)
        loss, parts = layer(sample, tg)
        loss.backward()
        rec = {"B": B, "C": C, "G": G, "scale": scale, "tseed": tseed, "loss": loss.detach().clone(),
               "parts": parts.clone(), "grad_abs_sum": float(sample.grad.double().abs().sum()),
               "grad_nnz": int((sample.grad != 0).sum())}
        if C <= 3:
            rec["grad"] = sample.grad.clone()
        layer.eval()
        det = layer(sample.detach())
        rec["det_rows"] = det[:, ::37].clone()
        yl[name] = rec
    out["yolo_layer"] = yl

    # (3) whole networks
    nets = {}
    for name, (cfg, S, B, C) in {"tiny_128": ("yolo_baseline_tiny.cfg", 128, 2, 1),
                                 "tiny_416": ("yolo_baseline_tiny.cfg", 416, 2, 1),
                                 "full_128": ("yolo_baseline.cfg", 128, 2, 1),
                                 "tiny_128_c80": ("yolo_baseline_tiny.cfg", 128, 2, 80)}.items():
        cfg_path = os.path.join(tmp, name + ".cfg")
        YO.write_cfg_copy(os.path.join(REF, "CVC-YOLOv3", "model_cfg", cfg), cfg_path, S, S, C, csv_path)
        torch.manual_seed(0)
        model = models.Darknet(cfg_path, 2.0, 1.6, 25.0, 0.1, True)
        model.apply(uu.weights_init_normal)
        model.train()
        x = YO.synth_images(B, S, S, seed=0)
        tg = YO.synth_targets(B, 16, seed=1)
        losses = model(x, tg)
        losses[0].sum().backward()
        rec = {"cfg": cfg, "S": S, "B": B, "C": C, "losses": torch.stack([l.detach() for l in losses]),
               "digest": param_digest(model.named_parameters()), "grads": sample_grads(model.named_parameters()),
               "running": {k: (float(v.double().sum()), float(v.double().abs().sum()))
                           for k, v in model.named_buffers() if "running" in k}}
        model.eval()
        with torch.no_grad():
            det = model(x)
        rec["det_shape"] = tuple(det.shape)
        rec["det_rows"] = det[:, ::53].clone()
        nets[name] = rec
        print(name, rec["losses"].tolist())
    out["darknet"] = nets


# ------------------------------------------------------------------------------------------ RektNet
def gen_rektnet(out):
    kn, crl = import_reference("RektNet", ["keypoint_net", "cross_ratio_loss"])
    res = {}
    B = 4
    x, thm, tpts = RO.synth_batch(B, seed=0)
    for loss_type in ("l2_softargmax", "l2_heatmap", "l1_softargmax"):
        for geo in (False, True):
            torch.manual_seed(17)
            net = kn.KeypointNet()
            net.train()
            hm, pts = net(x)
            loss_fn = crl.CrossRatioLoss(loss_type, geo, 0.055, 0.038)
            loc, g, total = loss_fn(hm, pts, thm, tpts)
            total.backward()
            res[f"{loss_type}_geo{int(geo)}"] = {
                "B": B, "loc": loc.detach().clone(), "geo": g.detach().clone().float(), "total": total.detach().clone(),
                "pts": pts.detach().clone(), "hm_rows": hm.detach()[:, :, ::16, ::16].clone(),
                "digest": param_digest(net.named_parameters()), "grads": sample_grads(net.named_parameters()),
                "running": {k: (float(v.double().sum()), float(v.double().abs().sum()))
                            for k, v in net.named_buffers() if "running" in k}}
            print(loss_type, geo, float(loc), float(g), float(total))
    torch.manual_seed(17)
    net = kn.KeypointNet()
    net.eval()
    with torch.no_grad():
        hm, pts = net(x)
    res["eval"] = {"pts": pts.clone(), "hm_rows": hm[:, :, ::16, ::16].clone()}
    out["rektnet"] = res


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    y, r = {}, {}
    gen_yolo(y)
    torch.save(y, os.path.join(OUT, "yolo_golden.pt"))
    gen_rektnet(r)
    torch.save(r, os.path.join(OUT, "rektnet_golden.pt"))
    for f in os.listdir(OUT):
        print(f, os.path.getsize(os.path.join(OUT, f)))
