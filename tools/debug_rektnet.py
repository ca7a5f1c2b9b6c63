"""RektNet B200 path vs oracle (fp32 and bf16-storage emulation): outputs and gradients (debug aid)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest  # noqa: F401,E402
from oracle import rektnet_oracle as RO  # noqa: E402

import cross_ratio_loss  # noqa: E402
import keypoint_net  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
loss_type = sys.argv[2] if len(sys.argv) > 2 else "l2_heatmap"
torch.manual_seed(17)
net = keypoint_net.KeypointNet()
x, thm, tpts = RO.synth_batch(B, seed=0)
res = {}
for emu in (False, True):
    params = {k: v.detach().clone().requires_grad_(True) for k, v in net.named_parameters()}
    buffers = {k: v.clone() for k, v in net.named_buffers()}
    hm, pts = RO.keypointnet_forward(params, buffers, x, True, emulate_bf16=emu)
    loc, geo, total = RO.cross_ratio_loss(hm, pts, thm, tpts, loss_type, True, 0.055, 0.038)
    total.backward()
    res[emu] = (hm.detach(), pts.detach(), float(loc), float(geo), {k: v.grad for k, v in params.items()})
net = net.cuda().train()
hm, pts = net(x.cuda())
loc, geo, total = cross_ratio_loss.CrossRatioLoss(loss_type, True, 0.055, 0.038)(hm, pts, thm.cuda(), tpts.cuda())
total.backward()
for emu in (False, True):
    rhm, rpts, rloc, rgeo, rg = res[emu]
    print(f"--- vs oracle emulate_bf16={emu}: loc {float(loc):.5f}/{rloc:.5f} geo {float(geo):.5f}/{rgeo:.5f} "
          f"pts maxerr {float((pts.detach().cpu() - rpts).abs().max()):.5f} hm relerr "
          f"{float((hm.detach().cpu() - rhm).norm() / rhm.norm()):.5f}")
    for k, p in net.named_parameters():
        r, g = rg[k], p.grad.cpu()
        if "bias" in k and "bn" not in k and k != "out.bias":
            continue
        cos = float((g * r).sum() / (g.norm() * r.norm() + 1e-30))
        print(f"   {k:28s} ref {float(r.norm()):.4e} got {float(g.norm()):.4e} cos {cos:.4f}")
