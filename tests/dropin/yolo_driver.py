"""Runs the UNMODIFIED reference CVC-YOLOv3/train.py (from baseline/_ref, staged by __graft_entry__.build()) on top of
the B200 drop-in modules: `models`, `utils.parse_config` and the hot-path functions of `utils.utils` come from
mit-driverless-cv-traininginfra_b200/CVC-YOLOv3, everything else (train.py, validate.py, utils/datasets.py, utils/nms.py,
the rest of utils/utils.py) is the reference's own file.  Only the DATA is synthetic: train.ImageLabelDataset is replaced
by an in-memory dataset with the reference's item format (uri, img[3,H,W] in [0,1], labels[T,5]) -- the reference's
loader needs its GCS-hosted images and imgaug/accimage.  Prints one JSON line.

usage: python yolo_driver.py <repo root> <work dir> [reference <start.weights>]
(cwd must be <work dir>: train.py writes logs/result.txt).  With `reference` the SAME driver runs the reference's own
models.py instead (on the CPU: CUDA_VISIBLE_DEVICES is emptied by the caller) from the given start weights, so the two
trajectories can be compared."""
import json
import os
import sys
import types

root, work = sys.argv[1], sys.argv[2]
reference_mode = len(sys.argv) > 3 and sys.argv[3] == "reference"
pkg = os.path.join(root, "mit-driverless-cv-traininginfra_b200")
ref = os.path.join(root, "baseline", "_ref")
if reference_mode:
    sys.path[:0] = [os.path.join(ref, "CVC-YOLOv3"), pkg, root]
else:
    os.environ["B200CV_REFERENCE_ROOT"] = ref  # utils/__init__.py + utils/utils.py resolve the other helpers there
    sys.path[:0] = [os.path.join(pkg, "CVC-YOLOv3"), os.path.join(ref, "CVC-YOLOv3"), pkg, root]

# third-party modules of the reference's data pipeline that are not installed here (never called: the data is synthetic)
for name in ("imgaug", "imgaug.augmenters", "tensorboardX"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["imgaug"].augmenters = sys.modules["imgaug.augmenters"]

import numpy as np  # noqa: E402
import torch  # noqa: E402
from PIL import Image  # noqa: E402

import train  # noqa: E402  -- the reference's script, byte for byte
import models  # noqa: E402  -- the B200 drop-in

assert os.path.realpath(train.__file__).startswith(os.path.realpath(ref)), train.__file__
assert os.path.realpath(models.__file__).startswith(os.path.realpath(ref if reference_mode else pkg)), models.__file__
assert train.Darknet is models.Darknet

from b200cv import cfg_gen, synth  # noqa: E402

S, B, N = 128, 4, 12
cfg1 = cfg_gen.write_cfg(work, "tiny", S, S, 1)
# validate.py:81-96 indexes a CPU tensor with CUDA indices when an image has NO detection above conf_thresh (a bug of
# the reference on a GPU with current torch); the shipped 0.8 is never reached by a 6-step model, so lower it
_text = open(cfg1).read().replace("conf_thresh=0.8", "conf_thresh=0.3")
open(cfg1, "w").write(_text)
img_path = os.path.join(work, "frame.png")
Image.fromarray((np.random.RandomState(0).rand(S, S, 3) * 255).astype(np.uint8)).save(img_path)


class SyntheticConeDataset(torch.utils.data.Dataset):
    """(uri, img, labels) items like utils/datasets.py:315."""

    def __init__(self, path, **kwargs):
        self.kw = kwargs
        self.imgs = synth.synth_images(N, S, S, seed=3)
        self.labels = synth.synth_targets(N, 16, seed=4)

    def __len__(self):
        return N

    def __getitem__(self, i):
        return img_path, self.imgs[i], self.labels[i]


train.ImageLabelDataset = SyntheticConeDataset
train.num_cpu = 0  # DataLoader workers: none needed for in-memory data

# start weights with 255-filter heads (the reference starts from COCO weights and keeps the first 18 filters)
if reference_mode:
    w0 = sys.argv[4]
else:
    torch.manual_seed(0)
    seed_model = models.Darknet(cfg_gen.write_cfg(work, "tiny", S, S, 80), 2.0, 1.6, 25.0, 0.1, True)
    seed_model.apply(__import__("utils.utils", fromlist=["weights_init_normal"]).weights_init_normal)
    w0 = os.path.join(work, "start.weights")
    seed_model.save_weights(w0)

os.makedirs(os.path.join(work, "logs"), exist_ok=True)
out_dir = os.path.join(work, "out")
os.makedirs(out_dir, exist_ok=True)
val_loss = train.main(evaluate=False, batch_size=B, optimizer_pick="Adam", model_cfg=cfg1, weights_path=w0,
                      output_path=out_dir, dataset_path=work, num_epochs=2, num_steps=100, checkpoint_interval=1,
                      augment_affine=False, augment_hsv=False, lr_flip=False, ud_flip=False, momentum=0.9, gamma=0.95,
                      lr=1e-4, weight_decay=0.0, vis_batch=False, data_aug=False, blur=False, salt=False, noise=False,
                      contrast=False, sharpen=False, ts=False, debug_mode=False, upload_dataset=False, xy_loss=2.0,
                      wh_loss=1.6, no_object_loss=25.0, object_loss=0.1, vanilla_anchor=True, val_tolerance=3,
                      min_epochs=3)

# run_epoch (train.py:49-93) once more by hand and compare with a plain loop over the same batches
torch.manual_seed(1)
net = models.Darknet(cfg1, 2.0, 1.6, 25.0, 0.1, True)
net.load_weights(os.path.join(out_dir, "2.weights"), [18, 18])  # a checkpoint of THIS model: 3 * (5 + 1) head filters
net = net.to(train.device)
loader = torch.utils.data.DataLoader(SyntheticConeDataset(""), batch_size=B, shuffle=False)
opt = torch.optim.SGD(net.parameters(), lr=1e-3, momentum=0.9)
net.train()
losses, _, _ = train.run_epoch("train", loader, 3, opt, net, 1, 1, [0])
launches = 0
if not reference_mode:
    torch.cuda.synchronize()
    launches = __import__("b200cv.lib", fromlist=["lib"]).lib().launches
print(json.dumps({"val_loss": float(val_loss), "files": sorted(os.listdir(out_dir)),
                  "result_txt": float(open(os.path.join(work, "logs", "result.txt")).read()),
                  "epoch_losses": [float(v) for v in losses], "gpu_launches": launches}))
