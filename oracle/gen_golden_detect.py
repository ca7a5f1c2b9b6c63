"""Generates tests/golden/detect_golden.pt: outputs of the REFERENCE's utils/nms.py (imported from /root/reference)
and of the cv2 of this image (the reference's `cv2.resize` pre-processing, RektNet/utils.py:73-76) on seeded inputs.
Run in the build container only; the fixture and this script are committed.

    python oracle/gen_golden_detect.py
"""
import importlib.util
import os
import sys

import cv2
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("B200CV_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)

from oracle import detect_oracle as DO  # noqa: E402


def ref_nms():
    spec = importlib.util.spec_from_file_location("_ref_nms", os.path.join(REF, "CVC-YOLOv3", "utils", "nms.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.nms


def reference_detect(nms, det, conf_thres, nms_thres):
    """detect.py:84-90 verbatim in behaviour (per image): filter, corners, nms -> (rows, boxes, scores)."""
    sel = det[:, 4] > conf_thres
    rows = torch.nonzero(sel).flatten()
    detections = det[sel]
    box_corner = torch.zeros((detections.shape[0], 4))
    xy = detections[:, 0:2]
    wh = detections[:, 2:4] / 2
    box_corner[:, 0:2] = xy - wh
    box_corner[:, 2:4] = xy + wh
    probabilities = detections[:, 4]
    keep = nms(box_corner, probabilities, nms_thres)
    return rows[keep], box_corner[keep], probabilities[keep]


def ref_utils():
    """The reference's utils/utils.py (average_precision, compute_ap, bbox_iou, xywh2xyxy) with its heavy, unrelated
    imports stubbed when they are missing here."""
    import types

    for name in ("imgaug", "imgaug.augmenters", "tqdm", "matplotlib", "matplotlib.pyplot", "cv2_unused"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    spec = importlib.util.spec_from_file_location("_ref_utils", os.path.join(REF, "CVC-YOLOv3", "utils", "utils.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def reference_image_ap(U, box_corner, probabilities, labels, width, height, iou_thres):
    """validate.py:95-130 for one image (the loop body is inline in validate(), which needs a dataloader and image
    files, so its statements are repeated here around the REFERENCE's own bbox_iou / xywh2xyxy / average_precision)."""
    if box_corner.shape[0] == 0:
        return None
    _, inds = torch.sort(-probabilities)
    box_corner, probabilities = box_corner[inds], probabilities[inds]
    labels = labels[(labels[:, 1:5] <= 0).sum(dim=1) == 0]
    target_boxes = U.xywh2xyxy(labels[:, 1:5])
    target_boxes[:, (0, 2)] *= width
    target_boxes[:, (1, 3)] *= height
    detected = torch.zeros(target_boxes.shape[0], dtype=torch.uint8)
    correct = torch.zeros(box_corner.shape[0], dtype=torch.uint8)
    ious = U.bbox_iou(box_corner.unsqueeze(1).expand(-1, target_boxes.shape[0], -1),
                      target_boxes.unsqueeze(0).expand(box_corner.shape[0], -1, -1))
    if [] in ious.data.tolist():
        return None
    best_is = torch.argmax(ious, dim=1)
    for i, iou in enumerate(ious):
        best_i = best_is[i]
        if ious[i, best_i] > iou_thres and detected[best_i] == 0:
            correct[i] = 1
            detected[best_i] = 1
    ap, r, p = U.average_precision(tp=correct, conf=probabilities, n_gt=labels.shape[0])
    return float(ap), float(r), float(p), correct


def main():
    nms = ref_nms()
    out = {"nms": {}, "resize": {}, "ap": {}, "cv2_version": cv2.__version__}
    # ---- NMS: cases WITHOUT tied scores (the reference's sort is unstable; ties are covered oracle-vs-kernel only)
    cases = {
        "c1_416": dict(B=4, rows=10647, C=1, seed=1, conf=0.8, nms=0.25, hot=24),
        "c80_tiny": dict(B=3, rows=2535, C=80, seed=2, conf=0.8, nms=0.25, hot=40),
        "many": dict(B=2, rows=4000, C=1, seed=3, conf=0.3, nms=0.5, hot=300),  # > top_k candidates
        "none": dict(B=2, rows=500, C=1, seed=4, conf=0.999999, nms=0.25, hot=0),
    }
    for name, c in cases.items():
        det = DO.synth_detections(c["B"], c["rows"], c["C"], seed=c["seed"], hot=c["hot"], ties=False)
        res = []
        for b in range(c["B"]):
            cand = det[b, :, 4][det[b, :, 4] > c["conf"]]
            assert cand.unique().numel() == cand.numel(), "tied candidate scores in a golden case"
            rows, boxes, scores = reference_detect(nms, det[b], c["conf"], c["nms"])
            res.append({"rows": rows.clone(), "boxes": boxes.clone(), "scores": scores.clone()})
        out["nms"][name] = dict(c, out=res)
        print(name, [int(r["rows"].numel()) for r in res])
    # ---- per-image AP on the reference's NMS output (untied candidate scores)
    U = ref_utils()
    for name, c in {"ap_416": dict(B=10, rows=10647, C=1, seed=11, conf=0.8, nms=0.25, iou=0.5, hot=24, T=16),
                    "ap_many": dict(B=5, rows=4000, C=1, seed=12, conf=0.3, nms=0.5, iou=0.3, hot=300, T=64)}.items():
        while True:  # a seed without tied candidate scores (the reference's sorts are unstable)
            det = DO.synth_detections(c["B"], c["rows"], c["C"], seed=c["seed"], hot=c["hot"], ties=False)
            cand = [det[b, :, 4][det[b, :, 4] > c["conf"]] for b in range(c["B"])]
            if all(v.unique().numel() == v.numel() for v in cand):
                break
            c["seed"] += 100
        labels = DO.synth_labels_for(det, c["B"], c["T"], c["conf"], seed=c["seed"])
        res = []
        for b in range(c["B"]):
            rows, boxes, scores = reference_detect(nms, det[b], c["conf"], c["nms"])
            assert scores.unique().numel() == scores.numel()
            res.append(reference_image_ap(U, boxes, scores, labels[b], 416, 416, c["iou"]))
        out["ap"][name] = dict(c, out=res)
        print(name, [None if r is None else (round(r[0], 4), round(r[1], 3), round(r[2], 3)) for r in res])
    # ---- cv2.resize on seeded crops (incl. identity, exact 2x decimation, 1-pixel-wide, up- and down-scaling)
    rng = np.random.RandomState(7)
    shapes = [(80, 80), (160, 160), (160, 100), (37, 23), (23, 61), (1, 50), (50, 1), (2, 2), (200, 131), (81, 79),
              (300, 17), (12, 240)]
    for i, (h, w) in enumerate(shapes):
        img = rng.randint(0, 256, size=(h, w, 3)).astype(np.uint8)
        out["resize"][f"r{i}_{h}x{w}"] = {"img": torch.from_numpy(img), "out": torch.from_numpy(cv2.resize(img, (80, 80)))}
    img = rng.randint(0, 256, size=(60, 90, 3)).astype(np.uint8)
    out["resize"]["nonsquare_64x48"] = {"img": torch.from_numpy(img), "out": torch.from_numpy(cv2.resize(img, (64, 48))),
                                        "size": (64, 48)}
    # ---- letterbox: torchvision pad(127) + resize (PIL BILINEAR) + to_tensor, exactly detect.py:62-72
    from PIL import Image
    import PIL
    import torchvision
    import torchvision.transforms.functional as TF

    out["letterbox"] = {"pil_version": PIL.__version__, "torchvision_version": torchvision.__version__, "cases": {}}
    for name, (h, w, S) in {"wide_down": (90, 160, 64), "tall_down": (150, 70, 64), "wide_up": (30, 48, 64),
                            "square_same": (64, 64, 64), "wide_to_96": (120, 200, 96), "odd": (77, 131, 64)}.items():
        img = rng.randint(0, 256, size=(h, w, 3)).astype(np.uint8)
        pad_h, pad_w, ratio = DO.calculate_padding(h, w, S, S)
        pil = TF.pad(Image.fromarray(img), padding=(pad_w, pad_h, pad_w, pad_h), fill=(127, 127, 127),
                     padding_mode="constant")
        ref = TF.to_tensor(TF.resize(pil, (S, S)))
        out["letterbox"]["cases"][name] = {"img": torch.from_numpy(img), "S": S, "out": ref.clone(),
                                           "geom": (ratio, pad_w, pad_h)}
    path = os.path.join(ROOT, "tests", "golden", "detect_golden.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
