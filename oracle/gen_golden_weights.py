"""Golden Darknet `.weights` files WRITTEN BY THE REFERENCE (CVC-YOLOv3/models.py:339-422), for the a-9 / f-3 parity
tests (tests/test_host_logic.py::test_weights_files_written_by_the_reference).  Test infrastructure only.

Run in the build container (imports /root/reference, CPU only):

    python oracle/gen_golden_weights.py

Recipe (a two-head mini network so the fixtures stay small; the byte format does not depend on the layer sizes):
  F0  seed file: random parameters + random BN running statistics of the 80-class model, written by OUR save_weights
  A   = reference(80 classes).load_weights(F0, [255, 255]); .seen = 777; reference.save_weights(A)   -- 255-filter heads
  B   = reference(1 class).load_weights(A, [255, 255])  (keeps the first 18 of the 255 head filters, :380-394);
        .seen = 778; reference.save_weights(B)
Stored: A, B (bytes) and the digests of the reference's parameters/buffers after each load.  The tests load A with
the product models.Darknet (80 and 1 classes), compare every tensor digest with the reference's, and require the
product's save_weights to reproduce A and B byte for byte.
"""
import os
import sys
import tempfile
import warnings

import torch

warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")
PKG = os.path.join(ROOT, "mit-driverless-cv-traininginfra_b200")
sys.path.insert(0, ROOT)

from oracle.gen_golden import import_reference, param_digest  # noqa: E402

MINI_LAYERS = """
[convolutional]
filters=8
size=3
stride=1

[maxpool]
size=2
stride=2

[convolutional]
filters=16
size=3
stride=1

[convolutional]
size=1
stride=1
filters=preyolo
activation=linear

[yolo]
note=head

[route]
layers = -3

[convolutional]
filters=8
size=1
stride=1

[upsample]
stride=2

[route]
layers = -1, 0

[convolutional]
filters=16
size=3
stride=1

[convolutional]
size=1
stride=1
filters=preyolo
activation=linear

[yolo]
note=head
"""


def mini_cfg(directory, classes):
    """cfg of the mini network (+ train.csv with the anchors row); same [net] keys as the shipped cfgs."""
    sys.path.insert(0, PKG)
    from b200cv import cfg_gen

    os.makedirs(directory, exist_ok=True)
    csv_path = os.path.join(directory, "train.csv")
    cfg_gen.write_anchor_csv(csv_path)
    text = cfg_gen._net(64, 64, classes, "3,4,5|0,1,2", "2,1", csv_path, "255,255") + MINI_LAYERS
    path = os.path.join(directory, f"mini_c{classes}.cfg")
    with open(path, "w") as f:
        f.write(text)
    return path


def state_digest(model):
    return param_digest(list(model.named_parameters()) + [(k, v) for k, v in model.named_buffers()
                                                          if v.dtype.is_floating_point])


def main():
    d = tempfile.mkdtemp()
    cfg80, cfg1 = mini_cfg(d, 80), mini_cfg(d, 1)
    # ---- seed file written by the product's save_weights (CPU: plain nn modules, no kernels involved)
    for p in (PKG, os.path.join(PKG, "CVC-YOLOv3")):
        sys.path.insert(0, p)
    import models as product_models

    torch.manual_seed(11)
    ours = product_models.Darknet(cfg80, 2.0, 1.6, 25.0, 0.1, True)
    g = torch.Generator().manual_seed(12)
    with torch.no_grad():
        for p in ours.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * 0.1)
        for k, b in ours.named_buffers():
            if k.endswith("running_mean"):
                b.copy_(torch.randn(b.shape, generator=g))
            elif k.endswith("running_var"):
                b.copy_(torch.rand(b.shape, generator=g) + 0.5)
    f0 = os.path.join(d, "seed.weights")
    ours.save_weights(f0)
    for k in [k for k in sys.modules if k == "models" or k == "utils" or k.startswith("utils.")]:
        sys.modules.pop(k)
    for p in (PKG, os.path.join(PKG, "CVC-YOLOv3")):
        sys.path.remove(p)

    # ---- the reference reads and re-writes it
    (ref_models,) = import_reference("CVC-YOLOv3", ["models"])
    r80 = ref_models.Darknet(cfg80, 2.0, 1.6, 25.0, 0.1, True)
    r80.load_weights(f0, [255, 255])
    r80.seen = 777
    a_path = os.path.join(OUT, "mini_c80_ref.weights")
    r80.save_weights(a_path)
    r1 = ref_models.Darknet(cfg1, 2.0, 1.6, 25.0, 0.1, True)
    r1.load_weights(a_path, [255, 255])
    seen_after_load = int(r1.seen)
    r1.seen = 778
    b_path = os.path.join(OUT, "mini_c1_ref.weights")
    r1.save_weights(b_path)
    torch.save({"digest_c80": state_digest(r80), "digest_c1": state_digest(r1), "seen_after_load": seen_after_load,
                "layers": MINI_LAYERS}, os.path.join(OUT, "weights_golden.pt"))
    print("wrote", a_path, os.path.getsize(a_path), "bytes;", b_path, os.path.getsize(b_path), "bytes")


if __name__ == "__main__":
    main()
