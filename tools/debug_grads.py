"""Per-parameter gradient error of the B200 Darknet path vs the oracle on identical weights (debug aid)."""
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest  # noqa: F401,E402  (sets sys.path)
import helpers  # noqa: E402
from oracle import yolo_oracle as YO  # noqa: E402

cfg_name = sys.argv[1] if len(sys.argv) > 1 else "yolo_baseline_tiny.cfg"
S = int(sys.argv[2]) if len(sys.argv) > 2 else 128
B = int(sys.argv[3]) if len(sys.argv) > 3 else 2
d = tempfile.mkdtemp()
model, path = helpers.make_darknet(d, cfg_name, S, 1)
params = {k: v.detach().clone().requires_grad_(True) for k, v in model.named_parameters()}
buffers = {k: v.clone() for k, v in model.named_buffers()}
x, tg = YO.synth_images(B, S, S, seed=0), YO.synth_targets(B, 16, seed=1)
want = YO.darknet_forward(YO.NetSpec(path), params, buffers, x, tg)
want[0].backward()
model = model.cuda().train()
got = model(x.cuda(), tg.cuda())
got[0].sum().backward()
print("loss ref", [round(float(v), 5) for v in want])
print("loss got", [round(float(v), 5) for v in got])
for k, p in model.named_parameters():
    r = params[k].grad
    g = p.grad.cpu()
    err = float((g - r).norm() / (r.norm() + 1e-20))
    cos = float((g * r).sum() / (g.norm() * r.norm() + 1e-20))
    print(f"{k:45s} ref_norm {float(r.norm()):10.4e} got_norm {float(g.norm()):10.4e} rel_l2 {err:8.4f} cos {cos:7.4f}")
