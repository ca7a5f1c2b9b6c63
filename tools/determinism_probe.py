#!/usr/bin/env python
"""Run-to-run spread of the training forward on identical inputs (same process, same weights): the BN statistics are
accumulated with fp32 atomics, so sums differ in the last bits between runs; bf16 storage and (for tiny spatial sizes)
BatchNorm over a few dozen samples amplify that.  Prints the relative spread of the loss tuple over N repeats."""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "mit-driverless-cv-traininginfra_b200")
for p in (ROOT, PKG, os.path.join(PKG, "CVC-YOLOv3"), os.path.join(PKG, "RektNet")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402


def spread(rows):
    t = torch.tensor(rows, dtype=torch.float64)
    return ((t.max(0).values - t.min(0).values) / t.mean(0).abs().clamp_min(1e-12)).tolist()


def main():
    import contextlib

    import cross_ratio_loss
    import keypoint_net
    import models
    from b200cv import cfg_gen
    from b200cv import synth
    from utils.utils import weights_init_normal

    dev = torch.device("cuda:0")
    d = tempfile.mkdtemp()
    n = 6
    for kind, S, B in (("darknet53", 128, 2), ("darknet53", 416, 8), ("tiny", 416, 8)):
        torch.manual_seed(0)
        net = models.Darknet(cfg_gen.write_cfg(d, kind, S, S, 1), 2.0, 1.6, 25.0, 0.1, True)
        net.apply(weights_init_normal)
        net = net.to(dev).train()
        x, tg = synth.synth_images(B, S, S).to(dev), synth.synth_targets(B, 16).to(dev)
        rows = []
        with torch.no_grad():
            for _ in range(n):
                rows.append([float(v) for v in net(x, tg)])
        print(f"{kind} {S}x{S} bs{B}: rel spread of (total,x,y,w,h,obj,noobj) over {n} runs:",
              ["%.1e" % v for v in spread(rows)])
    torch.manual_seed(5)
    kp = keypoint_net.KeypointNet().to(dev).train()
    with contextlib.redirect_stdout(sys.stderr):
        loss_fn = cross_ratio_loss.CrossRatioLoss("l2_softargmax", True, 0.055, 0.038)
    x, thm, tpts = (t.to(dev) for t in synth.synth_keypoint_batch(8, seed=0))
    rows = []
    with torch.no_grad():
        for _ in range(n):
            hm, pts = kp(x)
            rows.append([float(v) for v in loss_fn(hm, pts, thm, tpts)])
    print(f"KeypointNet 80x80 bs8: rel spread of (loc,geo,total) over {n} runs:", ["%.1e" % v for v in spread(rows)])


if __name__ == "__main__":
    main()
