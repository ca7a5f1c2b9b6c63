#!/usr/bin/env python
"""How well-conditioned is the Darknet gradient w.r.t. bf16 storage, as a function of training progress?

At random init the no-object term dominates the loss: its upstream gradient is spatially uniform, BatchNorm backward
cancels it, and what is left is decided by LeakyReLU signs -- 0.4 % activation rounding then changes the DIRECTION of
early-layer gradients.  This probe trains on a fixed synthetic batch and, every few steps, compares the
bf16-mode gradient with the fp32-parity-mode gradient of the same weights (cosine per parameter)."""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "mit-driverless-cv-traininginfra_b200")
for p in (ROOT, PKG, os.path.join(PKG, "CVC-YOLOv3"), os.path.join(PKG, "RektNet")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402


def grads_of(model, x, tg, precision):
    model.engine().set_precision(precision)
    model.zero_grad(set_to_none=True)
    out = model(x, tg)
    out[0].backward()
    return [float(v) for v in out], {k: p.grad.detach().clone() for k, p in model.named_parameters()}


def main():
    import models
    from b200cv import cfg_gen, synth
    from utils.utils import weights_init_normal

    kind = sys.argv[1] if len(sys.argv) > 1 else "darknet53"
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    lr = float(sys.argv[4]) if len(sys.argv) > 4 else 1e-3
    dev = torch.device("cuda:0")
    d = tempfile.mkdtemp()
    torch.manual_seed(0)
    net = models.Darknet(cfg_gen.write_cfg(d, kind, S, S, 1), 2.0, 1.6, 25.0, 0.1, True)
    net.apply(weights_init_normal)
    net = net.to(dev).train()
    x, tg = synth.synth_images(B, S, S).to(dev), synth.synth_targets(B, 16).to(dev)
    opt = torch.optim.Adam(net.parameters(), lr=lr)
    os.environ["B200CV_CUDA_GRAPH"] = "0"
    step = 0
    for upto in (0, 5, 10, 20, 40, 80, 160):
        net.engine().set_precision("bf16")
        while step < upto:
            opt.zero_grad()
            net(x, tg)[0].backward()
            opt.step()
            step += 1
        state = {k: v.clone() for k, v in net.state_dict().items()}
        l16, g16 = grads_of(net, x, tg, "bf16")
        net.load_state_dict(state)  # the probe passes moved the running statistics
        l32, g32 = grads_of(net, x, tg, "fp32")
        net.load_state_dict(state)
        cos, ratio = [], []
        for k in g16:
            a, b = g16[k].double().flatten(), g32[k].double().flatten()
            if float(b.norm()) == 0:
                continue
            cos.append(float((a * b).sum() / (a.norm() * b.norm() + 1e-30)))
            ratio.append(abs(float(a.norm() / (b.norm() + 1e-30)) - 1))
        cos.sort()
        ratio.sort()
        print(f"step {step:4d} loss bf16 {l16[0]:9.4f} fp32 {l32[0]:9.4f} rel {abs(l16[0]-l32[0])/abs(l32[0]):.1e} | "
              f"parts max rel {max(abs(a-b)/max(abs(b),1e-3) for a, b in zip(l16[1:], l32[1:])):.1e} | "
              f"cos min {cos[0]:.4f} p10 {cos[len(cos)//10]:.4f} median {cos[len(cos)//2]:.4f} | "
              f"norm dev median {ratio[len(ratio)//2]:.3f} max {ratio[-1]:.3f}", flush=True)


if __name__ == "__main__":
    main()
