"""Pins oracle/detect_oracle.py (CPU): its NMS must reproduce what the REFERENCE's utils/nms.py produced
(tests/golden/detect_golden.pt, written by oracle/gen_golden_detect.py), its resize what cv2 produced."""
import os

import numpy as np
import pytest
import torch

from oracle import detect_oracle as DO

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def golden_detect():
    return torch.load(os.path.join(ROOT, "tests", "golden", "detect_golden.pt"), weights_only=False)


def test_nms_matches_reference_outputs(golden_detect):
    assert len(golden_detect["nms"]) >= 4
    for name, c in golden_detect["nms"].items():
        det = DO.synth_detections(c["B"], c["rows"], c["C"], seed=c["seed"], hot=c["hot"], ties=False)
        for b, ref in enumerate(c["out"]):
            rows, boxes, scores = DO.detect_nms(det[b], c["conf"], c["nms"])
            assert torch.equal(rows, ref["rows"]), (name, b)  # kept set AND order: bit-exact
            assert torch.equal(boxes, ref["boxes"]), (name, b)
            assert torch.equal(scores, ref["scores"]), (name, b)


def test_nms_tie_rule_and_nan():
    # equal scores: the later row is visited first (stable ascending sort read from the end)
    boxes = torch.tensor([[0, 0, 10, 10], [0, 0, 10, 10], [20, 20, 30, 30.0]])
    scores = torch.tensor([0.9, 0.9, 0.5])
    assert DO.nms(boxes, scores, 0.5).tolist() == [1, 2]
    # zero-area duplicates: IoU = 0/0 = NaN is dropped by `IoU.le(overlap)`
    boxes = torch.tensor([[5, 5, 5, 5], [5, 5, 5, 5.0]])
    assert DO.nms(boxes, torch.tensor([0.9, 0.8]), 0.5).tolist() == [0]
    assert DO.nms(torch.zeros(0, 4), torch.zeros(0), 0.5).numel() == 0


def test_resize_matches_cv2_golden(golden_detect):
    assert len(golden_detect["resize"]) >= 12
    for name, c in golden_detect["resize"].items():
        got = DO.resize_linear_u8(c["img"].numpy(), c.get("size", (80, 80)))
        assert np.array_equal(got, c["out"].numpy()), name


def test_resize_matches_live_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.RandomState(3)
    for _ in range(60):
        h, w = int(rng.randint(1, 200)), int(rng.randint(1, 200))
        img = rng.randint(0, 256, size=(h, w, 3)).astype(np.uint8)
        assert np.array_equal(DO.resize_linear_u8(img, (80, 80)), cv2.resize(img, (80, 80))), (h, w)


def test_crop_rect_is_inside_and_non_empty():
    rng = np.random.RandomState(0)
    for _ in range(200):
        box = rng.uniform(-100, 600, size=4).astype(np.float32)
        x0, y0, x1, y1 = DO.crop_rect(box, 0.325, 0.0, 140.0, 1280, 720)
        assert 0 <= x0 < x1 <= 1280 and 0 <= y0 < y1 <= 720
    assert DO.crop_rect(np.array([50.0, 140.0, 100.0, 210.0], np.float32), 0.5, 0.0, 140.0, 1280, 720) == \
        (100, 140, 200, 280)


def test_prep_crop_layout():
    frames = DO.synth_frames(1, 120, 160, seed=1)
    out = DO.prep_crop(frames[0], (10, 20, 90, 100))
    assert out.shape == (3, 80, 80) and out.dtype == np.float32
    assert np.array_equal(out, (frames[0][20:100, 10:90].transpose(2, 0, 1) / 255.0).astype(np.float32))  # identity size


def test_image_ap_matches_reference_functions(golden_detect):
    """AP / recall / precision / TP flags vs the reference's own bbox_iou + average_precision + compute_ap."""
    assert len(golden_detect["ap"]) >= 2
    n_valid = 0
    for name, c in golden_detect["ap"].items():
        det = DO.synth_detections(c["B"], c["rows"], c["C"], seed=c["seed"], hot=c["hot"], ties=False)
        labels = DO.synth_labels_for(det, c["B"], c["T"], c["conf"], seed=c["seed"])
        for b, ref in enumerate(c["out"]):
            _, boxes, scores = DO.detect_nms(det[b], c["conf"], c["nms"])
            got = DO.image_ap(boxes, scores, labels[b], 416, 416, c["iou"])
            assert (got is None) == (ref is None), (name, b)
            if ref is None:
                continue
            n_valid += 1
            assert torch.equal(got[3], ref[3]), (name, b)  # true-positive flags: exact
            assert got[:3] == ref[:3], (name, b)            # same torch ops, same floats
    assert n_valid >= 10


def test_compute_ap_hand_case():
    # TP, FP, TP with 4 labels: recall .25 .25 .5, precision 1 .5 .667 -> envelope 1, .667, .667 -> AP = .25 + .25*.667
    ap, r, p = DO.average_precision(torch.tensor([1, 0, 1], dtype=torch.uint8), torch.tensor([0.9, 0.8, 0.7]), 4)
    assert float(ap) == pytest.approx(0.25 + 0.25 * 2 / 3, rel=1e-6)
    assert float(r) == pytest.approx(0.5) and float(p) == pytest.approx(2 / 3)


REF = os.environ.get("B200CV_REFERENCE", "/root/reference")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "CVC-YOLOv3")), reason="reference tree not present (GPU box)")
def test_oracle_vs_live_reference_nms_and_ap():
    """Where the reference tree is mounted (the build container), compare the oracle with the reference's own nms() and
    average_precision() on fresh random cases (untied scores), beyond the committed goldens."""
    import importlib.util
    import sys
    import types

    def load(name, rel):
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF, "CVC-YOLOv3", rel))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        return m

    ref_nms = load("_ref_nms_live", "utils/nms.py").nms
    stubbed = []
    for name in ("imgaug", "imgaug.augmenters", "tqdm", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
                stubbed.append(name)
    try:
        U = load("_ref_utils_live", "utils/utils.py")
    finally:
        for name in stubbed:
            sys.modules.pop(name, None)
    g = torch.Generator().manual_seed(123)
    for case in range(25):
        n = int(torch.randint(1, 400, (1,), generator=g))
        ctr = torch.rand(n, 2, generator=g) * 100
        wh = 2 + torch.rand(n, 2, generator=g) * 40
        boxes = torch.cat([ctr - wh / 2, ctr + wh / 2], 1)
        scores = torch.rand(n, generator=g)
        if scores.unique().numel() != n:
            continue
        thr = float(torch.rand(1, generator=g)) * 0.8
        assert torch.equal(DO.nms(boxes, scores, thr, 200), ref_nms(boxes, scores, thr, 200)), case
        tp = (torch.rand(n, generator=g) > 0.5).to(torch.uint8)
        n_gt = int(torch.randint(1, 50, (1,), generator=g))
        got, want = DO.average_precision(tp, scores, n_gt), U.average_precision(tp, scores, n_gt)
        assert [float(v) for v in got] == [float(v) for v in want], case


def test_letterbox_matches_pil_torchvision_golden(golden_detect):
    cases = golden_detect["letterbox"]["cases"]
    assert len(cases) >= 6
    for name, c in cases.items():
        got, geom = DO.letterbox(c["img"].numpy(), c["S"], c["S"])
        assert np.array_equal(got, c["out"].numpy()), name
        assert geom == c["geom"], name


def test_letterbox_matches_live_pil():
    PIL = pytest.importorskip("PIL")
    TF = pytest.importorskip("torchvision.transforms.functional")
    from PIL import Image

    rng = np.random.RandomState(5)
    for h, w, S in [(72, 128, 52), (200, 90, 64), (33, 57, 96)]:
        img = rng.randint(0, 256, size=(h, w, 3)).astype(np.uint8)
        pad_h, pad_w, _ = DO.calculate_padding(h, w, S, S)
        pil = TF.pad(Image.fromarray(img), padding=(pad_w, pad_h, pad_w, pad_h), fill=(127, 127, 127))
        ref = TF.to_tensor(TF.resize(pil, (S, S))).numpy()
        assert np.array_equal(DO.letterbox(img, S, S)[0], ref), (h, w, S)


def test_product_resampling_tables_equal_the_oracle():
    """b200cv/preprocess.py computes the Pillow coefficient tables on the host (product code, vectorised); they must
    be identical to the oracle's loop restatement for up- and down-scaling, tiny and large axes."""
    from b200cv import preprocess as PP

    for a, b in [(1280, 416), (720, 416), (1280, 608), (300, 416), (37, 64), (162, 64), (100, 128), (999, 7), (5, 300)]:
        for x, y in zip(DO.pil_bilinear_coeffs(a, b), PP.bilinear_tables(a, b)):
            assert x.shape == y.shape and np.array_equal(x, y), (a, b)
    for hw in [(720, 1280), (1280, 720), (300, 500), (37, 100), (416, 416)]:
        assert PP.calculate_padding(*hw, 416, 416) == DO.calculate_padding(*hw, 416, 416)
