"""YOLO loss kernel (fused forward sums + dlogits) at the BASELINE shape: 416^2, bs64, C=80, three scales.
Prints achieved GB/s against algorithmic bytes = read of the 5 box/objectness logits per anchor cell (fp32)
+ write of all dlogit channels (bf16, 256 per pixel) -- see DESIGN.md."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest  # noqa
import torch
import models
from b200cv import yolo_ops
from b200cv import synth

dev = "cuda"
B, C, T = 64, 80, 16
tg = synth.synth_targets(B, T, seed=1).to(dev)
consts = (2.0, 1.6, 0.1, 25.0)
work = []
for G, mask in ((13, (6, 7, 8)), (26, (3, 4, 5)), (52, (0, 1, 2))):
    z = torch.randn(B, G, G, 256, device=dev)
    sa = yolo_ops.scaled_anchors([models.vanilla_anchor_list[i] for i in mask], 416 / G, dev)
    yt = yolo_ops.yolo_targets(tg, sa, G, G, 0.5)
    work.append((z, yt, torch.empty(B, G, G, 256, device=dev, dtype=torch.bfloat16), torch.zeros(6, dtype=torch.float64, device=dev),
                 torch.empty(B * G * G, 16, device=dev)))
g = torch.ones(1, device=dev)
def run():
    for z, yt, d, s, dc in work:  # loss sums + cell gradients in one pass, then the streaming expansion
        yolo_ops.yolo_loss_cells(z, False, yt, C, consts, sums=s, dcell=dc, gscale=g)
        yolo_ops.yolo_expand_dlogits(dc, d, 3, C)
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
tot = 0.0
n = 10
for _ in range(n):
    flush.zero_()  # evict L2 (126 MB) between timed iterations
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    tot += e0.elapsed_time(e1)
ms = tot / n
cells = B * 3 * (13 * 13 + 26 * 26 + 52 * 52)
pix = B * (13 * 13 + 26 * 26 + 52 * 52)
alg = cells * 5 * 4 + pix * 256 * 2 + 2 * pix * 16 * 4  # + compact cell gradients written and re-read
survey = 2 * cells * 85 * 4
print(f"yolo_loss fused fwd+bwd, 3 scales: {ms*1e3:.1f} us; algorithmic {alg/1e6:.1f} MB -> {alg/ms/1e6:.0f} GB/s; "
      f"SURVEY formula (2 x fp32 head) {survey/1e6:.1f} MB -> {survey/ms/1e6:.0f} GB/s")
