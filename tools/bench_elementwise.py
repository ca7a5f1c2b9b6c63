"""Achieved HBM GB/s of the BN / elementwise kernels at Darknet-53 bs64 shapes (tuning aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest  # noqa
import torch
from b200cv import ops

def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

dev = "cuda"
shapes = [(64, 416, 416, 32), (64, 208, 208, 64), (64, 104, 104, 128), (64, 52, 52, 256), (64, 52, 52, 128),
          (64, 26, 26, 512), (64, 13, 13, 1024)]
tot = {"apply": 0, "reduce": 0, "bwd_apply": 0}
for shp in shapes:
    n, h, w, c = shp
    y = torch.randn(shp, device=dev).to(torch.bfloat16)
    da = torch.randn(shp, device=dev).to(torch.bfloat16)
    f = lambda k: torch.rand(k, device=dev) + 0.5
    scale, shift, mean, rstd, coef = f(c), f(c), f(c), f(c), f(3 * c)
    out = torch.empty_like(y)
    elems = y.numel()
    t1 = timeit(lambda: ops.bn_apply_act(y, scale, shift, ops.ACT_LEAKY, 0.1, out=out))
    t2 = timeit(lambda: ops.bn_bwd_reduce(da, y, None, scale, shift, mean, rstd, ops.ACT_LEAKY, 0.1))
    t3 = timeit(lambda: ops.bn_bwd_apply(da, y, None, scale, shift, mean, rstd, coef, ops.ACT_LEAKY, 0.1, out=out))
    t4 = timeit(lambda: out.copy_(y))
    print(f"{str(shp):24s} apply {t1*1e3:7.1f}us {elems*4/t1/1e6:6.0f} GB/s | reduce {t2*1e3:7.1f}us {elems*4/t2/1e6:6.0f} GB/s | "
          f"bwd_apply {t3*1e3:7.1f}us {elems*6/t3/1e6:6.0f} GB/s | torch copy {elems*4/t4/1e6:6.0f} GB/s")
