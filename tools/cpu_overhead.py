"""How long does the host take to ENQUEUE one training step (no device sync) vs the device time?"""
import os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest  # noqa
import torch
import helpers
from b200cv import synth
d = tempfile.mkdtemp()
model, _ = helpers.make_darknet(d, "yolo_baseline.cfg", 416, 80)
model = model.cuda().train()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
x, tg = synth.synth_images(B, 416, 416).cuda(), synth.synth_targets(B, 16).cuda()
params = list(model.parameters())
def step():
    for p in params: p.grad = None
    l = model(x, tg); l[0].sum().backward()
for _ in range(3): step()
torch.cuda.synchronize()
for _ in range(3):
    t0 = time.perf_counter(); step(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"B={B} enqueue {1e3*(t1-t0):.1f} ms, total {1e3*(t2-t0):.1f} ms")
