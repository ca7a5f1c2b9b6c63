#!/usr/bin/env python
"""Time of one optimizer step over the Darknet-53 parameter set (61.9 M fp32 parameters, 222 tensors):
b200cv FusedAdam / FusedSGD (one launch) next to torch.optim.Adam / SGD (foreach, and torch's own fused Adam).
Algorithmic bytes: Adam 28 B/parameter (read p,g,m,v; write p,m,v), SGD+momentum 20 B/parameter.

    python tools/bench_optim.py [--iters 20]
"""
import argparse
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "mit-driverless-cv-traininginfra_b200")
for p in (ROOT, PKG, os.path.join(PKG, "CVC-YOLOv3"), os.path.join(PKG, "RektNet")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    import models
    from b200cv import cfg_gen
    from b200cv import optim as boptim
    from b200cv.lib import lib

    dev = torch.device("cuda:0")
    d = tempfile.mkdtemp()
    net = models.Darknet(cfg_gen.write_cfg(d, "darknet53", 416, 416, 80), 2.0, 1.6, 25.0, 0.1, True).to(dev)
    params = list(net.parameters())
    n = sum(p.numel() for p in params)
    arena = torch.randn(n, device=dev) * 1e-3  # gradients as views of one flat arena, like the engine's
    off = 0
    for p in params:
        p.grad = arena[off:off + p.numel()].view_as(p)
        off += p.numel()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > L2

    def time_opt(opt):
        for _ in range(3):
            opt.step()
        ts = []
        for _ in range(args.iters):
            flush.zero_()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            opt.step()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        return ts[len(ts) // 2]

    out = {"parameters": n, "tensors": len(params), "l2": "256 MB flush between iterations"}
    for name, mk, bpp in (
            ("b200cv.FusedAdam", lambda: boptim.FusedAdam(params, lr=1e-4, weight_decay=5e-4), 28),
            ("torch.optim.Adam(foreach)", lambda: torch.optim.Adam(params, lr=1e-4, weight_decay=5e-4, foreach=True), 28),
            ("torch.optim.Adam(fused)", lambda: torch.optim.Adam(params, lr=1e-4, weight_decay=5e-4, fused=True), 28),
            ("b200cv.FusedSGD", lambda: boptim.FusedSGD(params, lr=1e-4, momentum=0.9, weight_decay=5e-4), 20),
            ("torch.optim.SGD(foreach)", lambda: torch.optim.SGD(params, lr=1e-4, momentum=0.9, weight_decay=5e-4,
                                                                foreach=True), 20)):
        opt = mk()
        ms = time_opt(opt)
        out[name] = {"step_ms": round(ms, 4), "GB/s": round(bpp * n / ms / 1e6, 1)}
        if name.startswith("b200cv"):  # the launch alone (events around the ABI call), L2 flushed
            flush.zero_()
            kms = sum(lib().profile_step(opt.step).values())
            out[name].update({"kernel_ms": round(kms, 4), "kernel_GB/s": round(bpp * n / kms / 1e6, 1)})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
