"""GPU parity of the detect -> NMS -> crop -> RektNet joint through the C ABI: kept sets, boxes and
crop bytes are bit-exact against the reference-made goldens and the oracle."""
import os

import numpy as np
import pytest
import torch

import helpers
from b200cv import detect_ops, pipeline
from oracle import detect_oracle as DO

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def golden_detect():
    return torch.load(os.path.join(ROOT, "tests", "golden", "detect_golden.pt"), weights_only=False)


def _check_image(d, b, rows, boxes, scores, tag):
    n = int(d.counts[b])
    assert n == rows.numel(), (tag, n, rows.numel())
    assert torch.equal(d.rows[b, :n].cpu().long(), rows), tag
    assert torch.equal(d.boxes[b, :n].cpu(), boxes), tag
    assert torch.equal(d.scores[b, :n].cpu(), scores), tag
    assert bool((d.rows[b, n:] == -1).all()) and bool((d.boxes[b, n:] == 0).all()), tag


def test_nms_bit_exact_against_reference_goldens(golden_detect):
    for name, c in golden_detect["nms"].items():
        det = DO.synth_detections(c["B"], c["rows"], c["C"], seed=c["seed"], hot=c["hot"], ties=False)
        d = detect_ops.detect_nms(det.cuda(), c["conf"], c["nms"], 200)
        for b, ref in enumerate(c["out"]):
            _check_image(d, b, ref["rows"], ref["boxes"], ref["scores"], (name, b))


@pytest.mark.parametrize("B,rows,C,hot,conf,nmst,topk", [
    (8, 10647, 80, 60, 0.8, 0.25, 200),   # Darknet-53 416^2 head, with exactly tied scores
    (4, 22743, 1, 400, 0.5, 0.45, 200),   # 608^2, more candidates than top_k (radix select)
    (3, 2535, 1, 500, 0.0, 0.6, 512),     # every row a candidate, top_k at the kernel limit
    (2, 700, 3, 10, 0.8, 0.25, 7),        # tiny top_k
    (5, 33, 1, 33, 0.1, 0.0, 200),        # fewer rows than threads; overlap threshold 0
])
def test_nms_random_vs_oracle(B, rows, C, hot, conf, nmst, topk):
    det = DO.synth_detections(B, rows, C, seed=B * 1000 + rows, hot=hot, ties=True)
    d = detect_ops.detect_nms(det.cuda(), conf, nmst, topk)
    for b in range(B):
        r, bx, sc = DO.detect_nms(det[b], conf, nmst, topk)
        _check_image(d, b, r, bx, sc, (B, rows, b))


def test_nms_all_scores_tied_and_degenerate_boxes():
    det = torch.zeros(1, 300, 6)
    det[0, :, 0:2] = torch.arange(300).float().reshape(-1, 1) * 3.0  # a diagonal chain of overlapping boxes
    det[0, :, 2:4] = 10.0
    det[0, :, 4] = 0.9
    det[0, 100:110, 2:4] = 0.0  # zero-area boxes (IoU NaN against an identical zero-area box)
    det[0, 105, 0:2] = det[0, 104, 0:2]
    d = detect_ops.detect_nms(det.cuda(), 0.5, 0.3, 200)
    r, bx, sc = DO.detect_nms(det[0], 0.5, 0.3, 200)
    _check_image(d, 0, r, bx, sc, "tied")


def test_utils_nms_drop_in():
    from utils.nms import nms

    det = DO.synth_detections(1, 3000, 1, seed=5, hot=80, ties=True)[0]
    _, boxes, scores = DO.filter_and_corners(det, 0.75)
    keep = nms(boxes.cuda(), scores.cuda(), 0.3)
    assert keep.dtype == torch.int64
    assert torch.equal(keep.cpu(), DO.nms(boxes, scores, 0.3))
    assert nms(torch.zeros(0, 4).cuda(), torch.zeros(0).cuda()).numel() == 0


def test_crop_resize_bit_exact_against_cv2_goldens(golden_detect):
    """Each golden image is placed in a frame; the box selects exactly that image; output bytes == cv2's."""
    for name, c in golden_detect["resize"].items():
        img = c["img"]
        h, w, _ = img.shape
        frame = torch.randint(0, 256, (1, h + 9, w + 13, 3), dtype=torch.uint8)
        frame[0, 4:4 + h, 6:6 + w] = img
        d = detect_ops.Detections(torch.tensor([[[6.0, 4.0, 6.0 + w, 4.0 + h]]]).cuda(), torch.ones(1, 1).cuda(),
                                  torch.zeros(1, 1, dtype=torch.int32).cuda(),
                                  torch.ones(1, dtype=torch.int32).cuda(), 1)
        offsets, src = detect_ops.compact(d)
        assert offsets.tolist() == [0, 1]
        size = c.get("size", (80, 80))
        crops, rects = detect_ops.crop_resize(frame.cuda(), d, src, 1, torch.tensor([1.0, 0.0, 0.0]), size)
        assert rects.tolist() == [[6, 4, 6 + w, 4 + h]], name
        want = (c["out"].numpy().transpose(2, 0, 1) / 255.0).astype(np.float32)
        assert np.array_equal(crops[0].cpu().numpy(), want), name


def test_crop_resize_random_boxes_vs_oracle():
    B, H, W, K = 3, 360, 640, 16
    frames = DO.synth_frames(B, H, W, seed=2)
    g = torch.Generator().manual_seed(9)
    ratio, pad_w, pad_h = 416.0 / 640.0, 0.0, 140.0
    boxes = torch.zeros(B, K, 4)
    ctr = torch.rand(B, K, 2, generator=g) * 416
    wh = 4 + torch.rand(B, K, 2, generator=g) * 150
    boxes[..., 0:2] = ctr - wh / 2
    boxes[..., 2:4] = ctr + wh / 2
    boxes[0, 0] = torch.tensor([-50.0, -50.0, 700.0, 700.0])   # whole frame after clipping
    boxes[0, 1] = torch.tensor([500.0, 500.0, 600.0, 600.0])   # entirely outside
    counts = torch.tensor([K, 5, 0], dtype=torch.int32)
    d = detect_ops.Detections(boxes.cuda(), torch.ones(B, K).cuda(), torch.zeros(B, K, dtype=torch.int32).cuda(),
                              counts.cuda(), K)
    offsets, src = detect_ops.compact(d)
    assert offsets.tolist() == [0, K, K + 5, K + 5]
    n = K + 5
    geom = torch.tensor([[ratio, pad_w, pad_h]] * B)
    crops, rects = detect_ops.crop_resize(torch.from_numpy(frames).cuda(), d, src, n, geom)
    src_c = src[:n].cpu()
    for i in range(n):
        b, s = src_c[i].tolist()
        rect = DO.crop_rect(boxes[b, s].numpy(), ratio, pad_w, pad_h, W, H)
        assert rects[i].tolist() == list(rect), i
        assert np.array_equal(crops[i].cpu().numpy(), DO.prep_crop(frames[b], rect)), i


def test_pipeline_end_to_end_vs_oracle_composition(cfg_dir):
    """Darknet (tiny, eval) -> NMS -> crop -> KeypointNet on the device; every stage after the network output is
    checked against the oracle applied to the device's own detections / crops."""
    import keypoint_net

    model, _ = helpers.make_darknet(cfg_dir, "yolo_baseline_tiny.cfg", 416, 1)
    model = model.cuda().eval()
    torch.manual_seed(3)
    kp = keypoint_net.KeypointNet().cuda().eval()
    B, H, W = 4, 360, 640
    imgs = torch.rand(B, 3, 416, 416, generator=torch.Generator().manual_seed(0)).cuda()
    frames = torch.from_numpy(DO.synth_frames(B, H, W, seed=4)).cuda()
    geom = torch.tensor([416.0 / 640.0, 0.0, 140.0])
    with torch.no_grad():
        det = model(imgs)
    conf = float(det[..., 4].flatten().float().kthvalue(det[..., 4].numel() - 40).values)  # ~40 candidates in all
    pipe = pipeline.ConePipeline(model, kp, conf_thres=conf, nms_thres=0.25)
    out = pipe(imgs, frames, geom, keep_heatmaps=True)
    det_c = det.float().cpu()
    n = 0
    for b in range(B):
        r, bx, sc = DO.detect_nms(det_c[b], conf, 0.25, 200)
        _check_image(out.detections, b, r, bx, sc, b)
        assert int(out.offsets[b]) == n
        for s in range(r.numel()):
            rect = DO.crop_rect(bx[s].numpy(), float(geom[0]), 0.0, 140.0, W, H)
            assert out.rects[n].tolist() == list(rect)
            n += 1
    assert out.n_crops == n and n > 0
    assert out.points.shape == (n, 7, 2) and out.heatmaps.shape == (n, 7, 80, 80)
    # the keypoints are what KeypointNet gives for exactly these crops
    crops, _ = detect_ops.crop_resize(frames, out.detections, detect_ops.compact(out.detections)[1], n, geom)
    with torch.no_grad():
        hm, pts = kp(crops)
    assert torch.equal(pts, out.points)
    assert torch.allclose(out.heatmaps.sum((2, 3)), torch.ones(n, 7, device="cuda"), atol=1e-4)
    pf = out.points_in_frame()
    assert bool((pf[..., 0] >= out.rects[:, None, 0]).all()) and bool((pf[..., 0] <= out.rects[:, None, 2]).all())


def _check_metrics(m, d, det, labels, conf, nmst, iou, topk, tag):
    B = det.shape[0]
    aps = []
    for b in range(B):
        _, boxes, scores = DO.detect_nms(det[b], conf, nmst, topk)
        ref = DO.image_ap(boxes, scores, labels[b], 416, 416, iou)
        if ref is None:
            assert int(m.valid[b]) == 0 and float(m.ap[b]) == 0.0, (tag, b)
            assert not bool(m.correct[b].any()), (tag, b)
            continue
        n = int(d.counts[b])
        assert int(m.valid[b]) == 1, (tag, b)
        assert torch.equal(m.correct[b, :n].cpu(), ref[3]), (tag, b)   # matching: bit-exact
        assert float(m.ap[b]) == pytest.approx(ref[0], rel=1e-6, abs=1e-7), (tag, b)
        assert float(m.recall[b]) == pytest.approx(ref[1], rel=1e-6), (tag, b)
        assert float(m.precision[b]) == pytest.approx(ref[2], rel=1e-6), (tag, b)
        aps.append(ref[0])
    if aps:
        assert m.means()[0] == pytest.approx(sum(aps) / len(aps), rel=1e-5)


def test_match_ap_against_reference_goldens(golden_detect):
    for name, c in golden_detect["ap"].items():
        det = DO.synth_detections(c["B"], c["rows"], c["C"], seed=c["seed"], hot=c["hot"], ties=False)
        labels = DO.synth_labels_for(det, c["B"], c["T"], c["conf"], seed=c["seed"])
        d = detect_ops.detect_nms(det.cuda(), c["conf"], c["nms"], 200)
        m = detect_ops.match_ap(d, labels.cuda(), 416, 416, c["iou"])
        for b, ref in enumerate(c["out"]):
            if ref is None:
                assert int(m.valid[b]) == 0, (name, b)
                continue
            n = int(d.counts[b])
            assert int(m.valid[b]) == 1 and torch.equal(m.correct[b, :n].cpu(), ref[3]), (name, b)
            assert float(m.ap[b]) == pytest.approx(ref[0], rel=1e-6, abs=1e-7), (name, b)
            assert float(m.recall[b]) == pytest.approx(ref[1], rel=1e-6), (name, b)
            assert float(m.precision[b]) == pytest.approx(ref[2], rel=1e-6), (name, b)


@pytest.mark.parametrize("B,rows,hot,conf,nmst,iou,T,topk", [
    (12, 10647, 40, 0.8, 0.25, 0.5, 16, 200),
    (6, 2535, 300, 0.2, 0.6, 0.1, 128, 512),   # many detections per label: duplicates become false positives
    (5, 500, 0, 0.95, 0.25, 0.5, 8, 200),      # no detections at all
])
def test_match_ap_random_vs_oracle(B, rows, hot, conf, nmst, iou, T, topk):
    det = DO.synth_detections(B, rows, 1, seed=rows + B, hot=hot, ties=True)
    labels = DO.synth_labels_for(det, B, T, conf, seed=B)
    d = detect_ops.detect_nms(det.cuda(), conf, nmst, topk)
    m = detect_ops.match_ap(d, labels.cuda(), 416, 416, iou)
    _check_metrics(m, d, det, labels, conf, nmst, iou, topk, (B, rows))


def test_detection_metrics_composition(cfg_dir):
    """validate.py's batch body on the device vs the oracle applied to the device's own detections."""
    model, _ = helpers.make_darknet(cfg_dir, "yolo_baseline_tiny.cfg", 416, 1)
    model = model.cuda().eval()
    B = 6
    imgs = torch.rand(B, 3, 416, 416, generator=torch.Generator().manual_seed(1)).cuda()
    with torch.no_grad():
        det = model(imgs).float().cpu()
    conf = float(det[..., 4].flatten().kthvalue(det[..., 4].numel() - 60).values)
    labels = DO.synth_labels_for(det, B, 16, conf, seed=3)
    m = pipeline.detection_metrics(model, imgs, labels.cuda(), conf_thres=conf, nms_thres=0.25, iou_thres=0.5)
    d = detect_ops.detect_nms(det.cuda(), conf, 0.25, 200)
    _check_metrics(m, d, det, labels, conf, 0.25, 0.5, 200, "composition")
    assert int(m.valid.sum()) >= 3


def test_letterbox_bit_exact_against_pil_goldens(golden_detect):
    from b200cv import preprocess

    for name, c in golden_detect["letterbox"]["cases"].items():
        img = c["img"]
        lb = preprocess.Letterbox(img.shape[:2], (c["S"], c["S"]), "cuda")
        frames = torch.stack([img, img.flip(0)]).contiguous().cuda()   # second frame: upside-down copy
        got = lb(frames)
        assert torch.equal(got[0].cpu(), c["out"]), name
        want1, _ = DO.letterbox(img.flip(0).numpy().copy(), c["S"], c["S"])
        assert np.array_equal(got[1].cpu().numpy(), want1), name
        assert lb.geom.tolist() == pytest.approx(list(c["geom"])), name
        # BGR frames -> RGB planes
        rev = lb(frames.flip(-1).contiguous(), reverse_channels=True)
        assert torch.equal(rev, got), name


def test_letterbox_full_frame_vs_oracle_and_pipeline_from_frames(cfg_dir):
    """720x1280 camera frames -> 416x416 network input (3.08x anti-aliased reduction, 280-row 127 borders) vs the
    oracle; then the whole joint from raw frames."""
    import keypoint_net
    from b200cv import preprocess

    B, H, W = 3, 720, 1280
    frames_np = DO.synth_frames(B, H, W, seed=6)  # treated as BGR
    frames = torch.from_numpy(frames_np).cuda()
    lb = preprocess.Letterbox((H, W), (416, 416), "cuda")
    imgs = lb(frames, reverse_channels=True)
    for b in range(B):
        want, geom = DO.letterbox(np.ascontiguousarray(frames_np[b][..., ::-1]), 416, 416)
        assert np.array_equal(imgs[b].cpu().numpy(), want), b
    assert geom == (416 / 1280, 0, 280)
    model, _ = helpers.make_darknet(cfg_dir, "yolo_baseline_tiny.cfg", 416, 1)
    model = model.cuda().eval()
    kp = keypoint_net.KeypointNet().cuda().eval()
    with torch.no_grad():
        det = model(imgs)
    conf = float(det[..., 4].flatten().float().kthvalue(det[..., 4].numel() - 30).values)
    pipe = pipeline.ConePipeline(model, kp, conf_thres=conf, nms_thres=0.25)
    out = pipe.from_frames(frames)
    ref = pipe(imgs, frames, lb.geom)
    assert out.n_crops == ref.n_crops > 0
    assert torch.equal(out.rects, ref.rects) and torch.equal(out.points, ref.points)
