#!/usr/bin/env python
"""Latency of the inference joint (BASELINE config 5): YOLOv3 416x416 detect -> NMS -> batched crop -> RektNet 80x80
keypoints at batch 128 on one B200, p50 / p99 over timed iterations (CUDA events around each whole call, a device
synchronisation on both sides).  Synthetic data: random-init weights, random images, synthetic 1280x720 frames; the
confidence threshold is set to the quantile of the network's own confidences that lets `--boxes` boxes per image
through (random weights never reach the cfg's 0.8), stated in the output.

    python tools/bench_pipeline.py [--batch 128] [--iters 30] [--boxes 8] [--out profiles/pipeline_rNN.json]

One JSON line.  `resident` = inputs already in HBM; `e2e` = pinned-host images + frames copied inside the timed region
and the keypoints read back.  The stage split comes from one extra event-timed pass over every ABI call.
"""
import argparse
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "mit-driverless-cv-traininginfra_b200")
for p in (ROOT, PKG, os.path.join(PKG, "CVC-YOLOv3"), os.path.join(PKG, "RektNet")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402


def percentile(xs, q):
    xs = sorted(xs)
    i = min(len(xs) - 1, max(0, int(round(q / 100.0 * (len(xs) - 1)))))
    return xs[i]


def run(batch=128, iters=30, warmup=3, boxes=8.0, classes=80, frame=(720, 1280)):
    args = argparse.Namespace(batch=batch, iters=iters, warmup=warmup, boxes=boxes, classes=classes, frame=frame)
    import keypoint_net
    import models
    from b200cv import cfg_gen, pipeline
    from b200cv.lib import lib
    from b200cv import synth
    from utils.utils import weights_init_normal

    dev = torch.device("cuda", torch.cuda.current_device())
    B, (H, W) = args.batch, args.frame
    d = tempfile.mkdtemp()
    torch.manual_seed(0)
    net = models.Darknet(cfg_gen.write_cfg(d, "darknet53", 416, 416, args.classes), 2.0, 1.6, 25.0, 0.1, True)
    net.apply(weights_init_normal)
    net = net.to(dev).eval()
    kp = keypoint_net.KeypointNet().to(dev).eval()

    from b200cv.preprocess import Letterbox

    base = torch.from_numpy(synth.synth_frames(8, H, W, seed=0))
    frames_h = base.repeat((B + 7) // 8, 1, 1, 1)[:B].contiguous().pin_memory()
    frames = frames_h.to(dev)
    lb = Letterbox((H, W), (416, 416), dev)
    imgs = lb(frames, reverse_channels=True)  # every path sees the same network input: the letterboxed frames
    imgs_h = imgs.cpu().pin_memory()
    geom = lb.geom

    with torch.no_grad():
        det = net(imgs)
    conf = det[..., 4].flatten().float()
    k = max(1, int(args.boxes * B))
    thres = float(conf.kthvalue(conf.numel() - k).values)
    pipe = pipeline.ConePipeline(net, kp, conf_thres=thres, nms_thres=net.get_threshs()[1])

    def run_resident():
        return pipe(imgs, frames, geom)

    def run_e2e():
        out = pipe(imgs_h.to(dev, non_blocking=True), frames_h.to(dev, non_blocking=True), geom)
        return out.points.cpu(), out.rects.cpu(), out.detections.counts.cpu()

    def run_frames_resident():
        return pipe.from_frames(frames)

    def run_frames_e2e():  # raw camera frames from pinned host memory: letterbox on the device, keypoints read back
        out = pipe.from_frames(frames_h.to(dev, non_blocking=True))
        return out.points.cpu(), out.rects.cpu(), out.detections.counts.cpu()

    res = {}
    for name, fn in (("resident", run_resident), ("e2e", run_e2e), ("frames_resident", run_frames_resident),
                     ("frames_e2e", run_frames_e2e)):
        for _ in range(args.warmup):
            fn()
        ts = []
        for _ in range(args.iters):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        res[name] = {"p50_ms": percentile(ts, 50), "p99_ms": percentile(ts, 99), "min_ms": min(ts), "max_ms": max(ts),
                     "images_per_s_at_p50": B / (percentile(ts, 50) / 1e3)}
    out = run_resident()
    stages = {"darknet": 0.0, "nms+compact": 0.0, "crop_resize": 0.0, "rektnet": 0.0}
    seen_crop = False
    detail = lib().profile_step(run_resident, detail=True)
    for name, _, ms in detail:  # launch order: Darknet calls, NMS, compact, crop, KeypointNet calls
        if name == "b200cv_crop_resize_u8":
            seen_crop = True
            stages["crop_resize"] += ms
        elif name in ("b200cv_detect_nms", "b200cv_detect_compact"):
            stages["nms+compact"] += ms
        else:
            stages["rektnet" if seen_crop else "darknet"] += ms
    line = {
        "metric": "latency ms per batch, detect -> NMS -> crop -> RektNet (BASELINE config 5)", "n_gpus": 1,
        "config": {"workload": f"Darknet-53 416x416 C={args.classes} eval bs{B} -> NMS top-200 -> crop+resize 80x80 "
                               f"from {W}x{H} BGR frames -> KeypointNet eval",
                   "network_input": "letterboxed synthetic frames (pad 127, PIL-bilinear resize)",
                   "conf_thres": thres, "nms_thres": pipe.nms_thres, "crops_per_batch": out.n_crops,
                   "crops_per_batch_from_frames": run_frames_resident().n_crops,
                   "candidate_boxes_per_image": args.boxes},
        "iters": args.iters, "warmup": args.warmup, "dtype": "bf16", "data": "synthetic",
        "resident": res["resident"], "e2e": dict(res["e2e"], h2d_bytes=imgs_h.numel() * 4 + frames_h.numel(),
                                                 d2h_bytes=out.n_crops * (14 * 4 + 16) + 4 * B),
        "from_frames": {"note": "network input made on the device by the letterbox kernel (pad 127, PIL-bilinear "
                                "resize, /255) from the same frames; e2e copies only the u8 frames",
                        "resident": res["frames_resident"],
                        "e2e": dict(res["frames_e2e"], h2d_bytes=frames_h.numel())},
        "stage_device_ms": {k_: round(v, 3) for k_, v in stages.items()},
    }
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--boxes", type=float, default=8.0, help="candidate boxes per image the threshold lets through")
    ap.add_argument("--classes", type=int, default=80)
    ap.add_argument("--frame", type=int, nargs=2, default=(720, 1280))
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    s = json.dumps(run(args.batch, args.iters, args.warmup, args.boxes, args.classes, tuple(args.frame)))
    print(s)
    if args.out:
        with open(os.path.join(ROOT, args.out), "w") as f:
            f.write(s + "\n")


if __name__ == "__main__":
    main()
