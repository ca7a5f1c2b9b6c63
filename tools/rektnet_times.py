"""Per-call device times of ONE eager RektNet 80x80 bs256 training step (l2_heatmap + geo), like layer_times.py."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest  # noqa: F401
import time
import torch
import cross_ratio_loss, keypoint_net
from b200cv.lib import lib
from b200cv import synth  # synthetic input recipe

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda")
torch.manual_seed(17)
net = keypoint_net.KeypointNet().to(dev).train()
x, thm, tpts = (t.to(dev) for t in synth.synth_keypoint_batch(B, seed=0))
loss_fn = cross_ratio_loss.CrossRatioLoss("l2_heatmap", True, 0.055, 0.038)

def step():
    for p in net.parameters():
        p.grad = None
    hm, pts = net(x)
    loss_fn(hm, pts, thm, tpts)[2].backward()

for _ in range(3):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    step()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / 10 * 1e3
os.environ["B200CV_CUDA_GRAPH"] = "0"
os.environ["B200CV_WGRAD_TAIL_FILL"] = "0"  # one stream: a kernel's own duration needs it to run alone  # the per-call timing needs the eager launches
step()
calls = lib().profile_step(step, detail=True)
tot = sum(ms for _, _, ms in calls)
print(f"# RektNet 80^2 bs{B}: wall {wall:.2f} ms/step, sum of per-call device times {tot:.2f} ms, {len(calls)} ABI calls")
agg = {}
for name, tag, ms in calls:
    a = agg.setdefault((name, tag), [0, 0.0]); a[0] += 1; a[1] += ms
for (name, tag), (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{name},{tag},{n},{ms:.3f}")
