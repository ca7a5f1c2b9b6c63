"""Shared helpers of the parity tests."""
import os

import torch

from b200cv import cfg_gen

REF_ROOT = os.environ.get("B200CV_REFERENCE", "/root/reference")
NET_KIND = {"yolo_baseline_tiny.cfg": "tiny", "yolo_baseline.cfg": "darknet53"}


def make_darknet(cfg_dir, cfg_name, S, C, consts=(2.0, 1.6, 25.0, 0.1), seed=0):
    """Product models.Darknet built with the golden recipe: manual_seed(0), weights_init_normal."""
    import models
    from utils.utils import weights_init_normal

    path = cfg_gen.write_cfg(cfg_dir, NET_KIND[cfg_name], S, S, C)
    torch.manual_seed(seed)
    model = models.Darknet(path, *consts, True)
    model.apply(weights_init_normal)
    return model, path


def digest(named):
    return {k: (float(v.detach().double().sum()), float(v.detach().double().abs().sum())) for k, v in named}


def assert_digest(named, gold, rel=1e-6):
    mine = digest(named)
    assert set(mine) == set(gold), (set(mine) ^ set(gold))
    for k, (s, a) in gold.items():
        assert abs(mine[k][0] - s) <= rel * (abs(a) + 1.0), k
        assert abs(mine[k][1] - a) <= rel * (abs(a) + 1.0), k


def grad_errors(named_params, gold_grads):
    """Per-parameter relative error of the sampled gradient values against the golden samples:
    max |g - g_ref| / (max|g_ref| over the sample + norm/sqrt(n))."""
    errs = {}
    for k, p in named_params:
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        ref = gold_grads[k]
        mine = g.detach().float().cpu().flatten()[ref["idx"]]
        scale = float(ref["val"].abs().max()) + ref["norm"] / max(1.0, p.numel() ** 0.5) + 1e-12
        errs[k] = float((mine - ref["val"]).abs().max()) / scale
    return errs


def write_mini_cfg(directory, classes, layers_text):
    """cfg of the two-head mini network of oracle/gen_golden_weights.py (same recipe, layer text from the golden)."""
    os.makedirs(directory, exist_ok=True)
    csv_path = os.path.join(directory, "train.csv")
    if not os.path.exists(csv_path):
        cfg_gen.write_anchor_csv(csv_path)
    path = os.path.join(directory, f"mini_c{classes}.cfg")
    with open(path, "w") as f:
        f.write(cfg_gen._net(64, 64, classes, "3,4,5|0,1,2", "2,1", csv_path, "255,255") + layers_text)
    return path
