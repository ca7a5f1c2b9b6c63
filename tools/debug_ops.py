"""Cross-checks every conv_wgrad / conv_dgrad / bn_bwd_apply call of one Darknet training step against
torch fp32 on the SAME device tensors (debug aid: isolates a broken kernel in realistic context)."""
import os
import sys
import tempfile

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest  # noqa: F401,E402
import helpers  # noqa: E402
from b200cv import ops  # noqa: E402
from oracle import yolo_oracle as YO  # noqa: E402

cfg_name = sys.argv[1] if len(sys.argv) > 1 else "yolo_baseline_tiny.cfg"
S = int(sys.argv[2]) if len(sys.argv) > 2 else 128
B = int(sys.argv[3]) if len(sys.argv) > 3 else 2

orig_wgrad, orig_dgrad, orig_apply = ops.conv_wgrad, ops.conv_dgrad, ops.bn_bwd_apply


def rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-20))


def wgrad(x, dy, cout, k, stride, pad, dil=1):
    out = orig_wgrad(x, dy, cout, k, stride, pad, dil)
    xr = x.float().permute(0, 3, 1, 2).contiguous().requires_grad_(False)
    dyr = dy.float()[..., :cout].permute(0, 3, 1, 2).contiguous()
    w = torch.zeros(cout, x.shape[-1], k, k, device=x.device, requires_grad=True)
    with torch.enable_grad():
        F.conv2d(xr, w, None, stride, pad, dil).backward(dyr)
    ref = w.grad.permute(0, 2, 3, 1).reshape(cout, k * k, x.shape[-1])
    print(f"  wgrad x{tuple(x.shape)} dy{tuple(dy.shape)} k{k} s{stride}: rel {rel(out, ref):.5f}")
    return out


def dgrad(dy, wpk_t, cin_fwd, k, stride, pad, dil, out_hw, out=None, residual=None):
    res_copy = residual.clone() if residual is not None else None
    o = orig_dgrad(dy, wpk_t, cin_fwd, k, stride, pad, dil, out_hw, out=out, residual=residual)
    # wpk_t: [I][RS][Opad] -> OIHW
    opad = wpk_t.shape[-1]
    w = wpk_t.float().view(cin_fwd, k, k, opad).permute(3, 0, 1, 2).contiguous()
    xz = torch.zeros(dy.shape[0], cin_fwd, out_hw[0], out_hw[1], device=dy.device, requires_grad=True)
    with torch.enable_grad():
        F.conv2d(xz, w, None, stride, pad, dil).backward(dy.float().permute(0, 3, 1, 2).contiguous())
    ref = xz.grad.permute(0, 2, 3, 1)
    if res_copy is not None:
        ref = ref + res_copy.float()[..., :cin_fwd]
    print(f"  dgrad dy{tuple(dy.shape)} -> {tuple(o.shape)} k{k} s{stride} res={residual is not None}: rel {rel(o.float()[..., :cin_fwd], ref):.5f}")
    return o


def apply(da, y, aout, scale, shift, mean, rstd, coef, act, slope, out=None):
    o = orig_apply(da, y, aout, scale, shift, mean, rstd, coef, act, slope, out=out)
    yf, daf = y.float(), da.float()
    z = yf * scale + shift
    d = torch.where(z > 0, torch.ones_like(z), torch.full_like(z, slope)) if act == 1 else torch.ones_like(z)
    dz = daf * d
    xh = (yf - mean) * rstd
    m = yf.numel() // yf.shape[-1]
    k1 = dz.sum((0, 1, 2)) / m
    k2 = (dz * xh).sum((0, 1, 2)) / m
    gamma_rstd = coef[:yf.shape[-1]]
    ref = gamma_rstd * (dz - k1 - xh * k2)
    print(f"  bn_bwd_apply {tuple(y.shape)}: rel {rel(o.float(), ref):.5f}  k1err {rel(coef[yf.shape[-1]:2*yf.shape[-1]], k1):.5f} k2err {rel(coef[2*yf.shape[-1]:], k2):.5f}")
    return o


ops.conv_wgrad, ops.conv_dgrad, ops.bn_bwd_apply = wgrad, dgrad, apply
d = tempfile.mkdtemp()
model, path = helpers.make_darknet(d, cfg_name, S, 1)
model = model.cuda().train()
x, tg = YO.synth_images(B, S, S, seed=0).cuda(), YO.synth_targets(B, 16, seed=1).cuda()
got = model(x, tg)
got[0].sum().backward()
torch.cuda.synchronize()
