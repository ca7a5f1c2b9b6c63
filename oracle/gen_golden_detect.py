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


def main():
    nms = ref_nms()
    out = {"nms": {}, "resize": {}, "cv2_version": cv2.__version__}
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
    path = os.path.join(ROOT, "tests", "golden", "detect_golden.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
