"""Runs the UNMODIFIED reference RektNet/train_eval.py `train_model` / `eval_model` (from baseline/_ref) on top of the
B200 drop-in keypoint_net / resnet / cross_ratio_loss modules, with an in-memory dataset in ConeDataset's item format
(img[3,80,80], hm[7,80,80], pts[7,2], name, shape) -- RektNet/dataset.py:56.  Prints one JSON line.

usage: python rektnet_driver.py <repo root> <work dir> [reference]
With `reference` the same loop runs on the reference's own keypoint_net / cross_ratio_loss (CPU: the caller empties
CUDA_VISIBLE_DEVICES), so the two training trajectories can be compared."""
import json
import os
import sys
import types

root, work = sys.argv[1], sys.argv[2]
reference_mode = len(sys.argv) > 3 and sys.argv[3] == "reference"
pkg = os.path.join(root, "mit-driverless-cv-traininginfra_b200")
ref = os.path.join(root, "baseline", "_ref")
sys.path[:0] = ([] if reference_mode else [os.path.join(pkg, "RektNet")]) + [os.path.join(ref, "RektNet"), pkg, root]

# RektNet/utils.py imports google.cloud.storage at module level (dataset download helper; not installed, never called)
g = types.ModuleType("google")
gc = types.ModuleType("google.cloud")
gcs = types.ModuleType("google.cloud.storage")
g.cloud, gc.storage = gc, gcs
sys.modules.update({"google": g, "google.cloud": gc, "google.cloud.storage": gcs})

import torch  # noqa: E402

import train_eval  # noqa: E402  -- the reference's script, byte for byte
import keypoint_net  # noqa: E402  -- the B200 drop-in

assert os.path.realpath(train_eval.__file__).startswith(os.path.realpath(ref)), train_eval.__file__
assert os.path.realpath(keypoint_net.__file__).startswith(os.path.realpath(ref if reference_mode else pkg))
assert train_eval.KeypointNet is keypoint_net.KeypointNet

from b200cv import synth  # noqa: E402


class SyntheticCones(torch.utils.data.Dataset):
    def __init__(self, n, seed):
        self.x, self.hm, self.pts = synth.synth_keypoint_batch(n, seed=seed)

    def __len__(self):
        return self.x.shape[0]

    def __getitem__(self, i):
        return self.x[i], self.hm[i], self.pts[i], f"cone_{i}.jpg", torch.tensor([80, 80])


torch.manual_seed(17)
model = train_eval.KeypointNet(7, (80, 80), onnx_mode=False).to(train_eval.device)
loss_fn = train_eval.CrossRatioLoss("l2_heatmap", True, 0.055, 0.038)
opt = torch.optim.Adam(model.parameters(), lr=1e-3)
sched = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=0.999)
train_loader = torch.utils.data.DataLoader(SyntheticCones(32, 0), batch_size=8, shuffle=False, num_workers=0)
val_loader = torch.utils.data.DataLoader(SyntheticCones(4, 1), batch_size=1, shuffle=False, num_workers=0)
before = train_eval.eval_model(model=model, dataloader=val_loader, loss_function=loss_fn, input_size=(80, 80))
train_eval.train_model(model=model, output_uri=work, dataloader=train_loader, loss_function=loss_fn, optimizer=opt,
                       scheduler=sched, epochs=3, val_dataloader=val_loader, intervals=1, input_size=(80, 80), num_kpt=7,
                       save_checkpoints=False, kpt_keys=None, study_name="dropin", evaluate_mode=False)
after = train_eval.eval_model(model=model, dataloader=val_loader, loss_function=loss_fn, input_size=(80, 80))
launches = 0
if not reference_mode:
    torch.cuda.synchronize()
    launches = __import__("b200cv.lib", fromlist=["lib"]).lib().launches
print(json.dumps({"val_before": [float(v) for v in before], "val_after": [float(v) for v in after],
                  "lr": opt.param_groups[0]["lr"], "gpu_launches": launches}))
