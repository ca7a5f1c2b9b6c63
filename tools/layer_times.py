"""Per-call device times of ONE eager Darknet-53 416^2 bs64 training step (CUDA events around every ABI call),
aggregated per (entry point, shape): conv calls with achieved TFLOP/s, BN/elementwise passes with GB/s.
Measurement aid (run on a B200): python tools/layer_times.py [size] [batch] > gpurun_out/layer_times.txt"""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest  # noqa: F401  (sets sys.path for the package)
import torch
import models
from b200cv import cfg_gen
from b200cv.lib import lib
from b200cv import synth  # synthetic input recipe
from utils.utils import weights_init_normal

size = int(sys.argv[1]) if len(sys.argv) > 1 else 416
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
os.environ["B200CV_CUDA_GRAPH"] = "0"
os.environ["B200CV_WGRAD_TAIL_FILL"] = "0"  # one stream: a kernel's own duration needs it to run alone
dev = torch.device("cuda")
cfg = cfg_gen.write_cfg(tempfile.mkdtemp(), "darknet53", size, size, 80)
torch.manual_seed(0)
model = models.Darknet(cfg, 2.0, 1.6, 25.0, 0.1, True)
model.apply(weights_init_normal)
model = model.to(dev).train()
x, t = synth.synth_images(B, size, size).to(dev), synth.synth_targets(B, 16).to(dev)

def step():
    for p in model.parameters():
        p.grad = None
    model(x, t)[0].sum().backward()

for _ in range(3):
    step()
runs = [lib().profile_step(step, detail=True) for _ in range(3)]
agg = {}
for calls in runs:
    for name, tag, ms in calls:
        k = (name, tag)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ms
n = len(runs)
tot = sum(v[1] for v in agg.values()) / n
print(f"# Darknet-53 {size}^2 bs{B}: sum of per-call device times {tot:.2f} ms/step (eager, events around each call)")
byname = {}
for (name, tag), (cnt, ms) in agg.items():
    byname[name] = byname.get(name, 0.0) + ms / n
for name, ms in sorted(byname.items(), key=lambda kv: -kv[1]):
    print(f"# {name:32s} {ms:8.3f} ms  {100*ms/tot:5.1f}%")
print("entry,shape,calls/step,ms/step,ms/call,TFLOP/s|GB/s")
for (name, tag), (cnt, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    per_step, per_call = ms / n, ms / cnt
    rate = ""
    if tag and name.startswith("b200cv_conv"):
        cin, cout, k, s, nb, oh, ow = tag
        fl = 2.0 * cin * cout * k * k * nb * oh * ow
        rate = f"{fl / per_call / 1e9:8.1f} TF"
    elif tag and len(tag) == 2:
        rows, c = tag
        nbytes = {"b200cv_bn_apply_act": 4, "b200cv_bn_bwd_reduce": 4, "b200cv_bn_bwd_apply": 6, "b200cv_copy_slice": 4}.get(name)
        if nbytes:
            rate = f"{rows * c * nbytes / per_call / 1e6:8.0f} GB/s"
    print(f"{name},{tag},{cnt // n},{per_step:.3f},{per_call:.4f},{rate}")
