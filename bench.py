#!/usr/bin/env python
"""Benchmark of the hot path: images/sec, forward+backward, CVC-YOLOv3 Darknet-53 416x416 bs64 per GPU
(BASELINE.json metric; RektNet 80x80 bs256 reported as a secondary figure), synthetic data.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One JSON line on rank 0.  `value` = whole-job img/s with inputs resident in HBM; `e2e` = the same step
through the public API with pinned-host inputs (H2D inside the timed region, loss read back);
`roofline` = the tcgen05 convolution kernels' achieved TFLOP/s (algorithmic conv FLOPs of a step / CUDA-event
time of the conv launches) against the measured sustained bf16 peak; `cpu_baseline` = the oracle (a port of
the reference's PyTorch CPU path) timed on this box's host cores on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "mit-driverless-cv-traininginfra_b200")
for p in (ROOT, PKG, os.path.join(PKG, "CVC-YOLOv3"), os.path.join(PKG, "RektNet")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

METRIC = "images/sec fwd+bwd (YOLOv3 Darknet-53 416x416 bs64 per GPU)"
IMG, BATCH, CLASSES, TMAX = 416, 64, 80, 16
LOSS_CONSTS = (2.0, 1.6, 25.0, 0.1)  # xy, wh, no_object, object (CVC-YOLOv3/train.py:312-315)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"bf16_sustained": d.get("bf16_tflops_sustained", 1400.0), "bf16_burst": d.get("bf16_tflops", 1590.0),
                "hbm": d.get("hbm_gbs", 6650.0), "src": "measured"}
    return {"bf16_sustained": 1400.0, "bf16_burst": 1590.0, "hbm": 6650.0, "src": "fallback"}


def ncu_traffic():
    """DRAM bytes of the conv launches of one step, from the committed ncu launch list (profiles/roofline_rNN.json,
    written by tools/make_profile_summary.py); None when no capture is committed."""
    best = None
    pdir = os.path.join(ROOT, "profiles")
    if os.path.isdir(pdir):
        for name in sorted(os.listdir(pdir)):
            if name.startswith("roofline_") and name.endswith(".json"):
                best = json.load(open(os.path.join(pdir, name)))
    return best


def conv_flops_per_image(spec_layers, size):
    """Algorithmic conv FLOPs per image (2*Cin*Cout*k*k*Hout*Wout), forward; and fwd+dgrad+wgrad."""
    h = size
    hs = []
    fwd = 0.0
    tot = 0.0
    first = True
    for L in spec_layers:
        t = L["type"]
        if t == "convolutional":
            ho = (h + 2 * L["pad"] - L["k"]) // L["stride"] + 1
            f = 2.0 * L["cin"] * L["cout"] * L["k"] * L["k"] * ho * ho
            fwd += f
            tot += f * (2 if first else 3)  # the first layer needs no data gradient
            first = False
            h = ho
        elif t == "maxpool":
            h = h // 2 if L["stride"] == 2 else h
        elif t == "upsample":
            h = h * 2
        elif t == "route":
            h = hs[L["layers"][0] if L["layers"][0] >= 0 else len(hs) + L["layers"][0]]
        elif t == "shortcut":
            h = hs[-1]
        hs.append(h)
    return fwd, tot


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.proc, self.path = None, None
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(gpu_index)], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for line in open(self.path).read().strip().split("\n"):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            sm.sort()
            out = {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        return out


# ----------------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    """The reference's own CPU implementation of the path = the oracle port (PyTorch fp32 on the host cores),
    same metric/config, each step a bounded sample (batch 8 of the bs64 workload)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from b200cv import cfg_gen
    from oracle import yolo_oracle as YO

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample_b = 8
    d = tempfile.mkdtemp()
    spec = YO.NetSpec(cfg_gen.write_cfg(d, "darknet53", IMG, IMG, CLASSES))
    params, buffers = YO.init_params(spec, seed=0)
    for p in params.values():
        p.requires_grad_(True)
    x, tg = YO.synth_images(sample_b, IMG, IMG), YO.synth_targets(sample_b, TMAX)

    def step():
        for p in params.values():
            p.grad = None
        out = YO.darknet_forward(spec, params, buffers, x, tg, LOSS_CONSTS)
        out[0].backward()

    for _ in range(max(1, min(args.warmup, 2))):
        step()
    k = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(k):
        step()
    dt = (time.perf_counter() - t0) / k
    v = sample_b / dt
    sample = f"Darknet-53 {IMG}x{IMG} C={CLASSES} fwd+bwd at batch {sample_b} (bounded sample of the bs{BATCH} step), {k} steps"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "img/s", "n_gpus": args.gpus, "steps": k,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"CVC-YOLOv3 Darknet-53 {IMG}x{IMG} bs{BATCH} fwd+bwd, classes={CLASSES}, T={TMAX}"},
        "cpu_baseline": {"value": v, "unit": "img/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ----------------------------------------------------------------------------------------------- B200 arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--ncu-window", action="store_true",
                    help="after warm-up run ONE eager step between cudaProfilerStart/Stop and exit "
                         "(for `ncu --profile-from-start off`; prints no bench line)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    from b200cv import cfg_gen, parallel
    from b200cv.lib import lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback")
    local = parallel.init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    rank, world = parallel.rank(), parallel.world_size()
    warmup = max(3, args.warmup)
    steps = max(1, args.steps)

    import models
    from utils.utils import weights_init_normal

    d = tempfile.mkdtemp()
    cfg = cfg_gen.write_cfg(d, "darknet53", IMG, IMG, CLASSES)
    torch.manual_seed(0)
    model = models.Darknet(cfg, *LOSS_CONSTS, True)
    model.apply(weights_init_normal)
    model = model.to(dev).train()
    params = list(model.parameters())

    from b200cv import synth  # synthetic inputs of the named shapes (host-side torch; the oracle is not involved)

    imgs_h = synth.synth_images(BATCH, IMG, IMG, seed=rank).pin_memory()
    tg_h = synth.synth_targets(BATCH, TMAX, seed=1 + rank).pin_memory()
    imgs_d, tg_d = imgs_h.to(dev), tg_h.to(dev)

    def step(x, t):
        for p in params:
            p.grad = None
        losses = model(x, t)
        losses[0].sum().backward()
        return losses

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = lib().launches
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t)
        return ms, lib().launches - l0

    if args.ncu_window:
        os.environ["B200CV_CUDA_GRAPH"] = "0"
    for _ in range(warmup):
        step(imgs_d, tg_d)
    if args.ncu_window:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step(imgs_d, tg_d)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    sampler = ClockSampler(local) if rank == 0 else None
    ms, launches = timed(lambda: step(imgs_d, tg_d), steps)
    clocks = sampler.stop() if sampler else None
    value = world * BATCH * steps / (ms / 1e3)

    # end to end through the public API: every step copies ITS inputs from pinned host memory (double-buffered on
    # a copy stream, so the copy of step i+1 overlaps the compute of step i) and reads the 7 losses back
    bufs = [(torch.empty_like(imgs_d), torch.empty_like(tg_d)) for _ in range(2)]
    copy_stream = torch.cuda.Stream()
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    freed = [torch.cuda.Event(), torch.cuda.Event()]

    def prefetch(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[slot])
            bufs[slot][0].copy_(imgs_h, non_blocking=True)
            bufs[slot][1].copy_(tg_h, non_blocking=True)
            ready[slot].record(copy_stream)

    def e2e_run(n):
        for ev in freed:
            ev.record()
        prefetch(0)
        for i in range(n):
            slot = i & 1
            if i + 1 < n:
                prefetch(slot ^ 1)
            torch.cuda.current_stream().wait_event(ready[slot])
            losses = step(*bufs[slot])
            freed[slot].record()
            host = torch.stack([l.detach() for l in losses]).cpu()  # device -> host read of the step's result
        return host

    e2e_run(2)
    ms_e2e, _ = timed(lambda: e2e_run(steps), 1)
    e2e_value = world * BATCH * steps / (ms_e2e / 1e3)

    out = None
    # roofline of the dominant kernels: event-time every conv launch of one more step (every rank runs the
    # step -- it contains the gradient all-reduce -- rank 0 reports)
    os.environ["B200CV_CUDA_GRAPH"] = "0"  # per-launch events need the eager launches, not the captured graph
    prof = lib().profile_step(lambda: step(imgs_d, tg_d))
    os.environ["B200CV_CUDA_GRAPH"] = "1"
    if rank == 0:
        fwd_f, tot_f = conv_flops_per_image(synth.conv_layer_table(model), IMG)
        pk = peaks()
        conv_ms = sum(v for k, v in prof.items() if k in ("b200cv_conv_fwd", "b200cv_conv_dgrad", "b200cv_conv_wgrad"))
        step_ms = sum(prof.values())
        achieved = tot_f * BATCH / (conv_ms / 1e3) / 1e12
        out = {
            "metric": METRIC, "value": value, "unit": "img/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"CVC-YOLOv3 Darknet-53 {IMG}x{IMG} bs{BATCH}/GPU fwd+bwd, classes={CLASSES}, T={TMAX}",
                       "parallelism": f"dp{world}", "global_batch": BATCH * world,
                       "l2": "activations per step (>10 GB) exceed the 126 MB L2; no explicit flush"},
            "e2e": {"value": e2e_value, "unit": "img/s", "h2d_bytes_per_step": imgs_h.numel() * 4 + tg_h.numel() * 4,
                    "d2h_bytes_per_step": 28, "ms_per_step": ms_e2e / steps},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                         "frac": achieved / pk["bf16_sustained"],
                         "traffic": (ncu_traffic() or {}).get("conv_dram_bytes_per_step"),
                         "traffic_note": "DRAM bytes (read+write) of all conv launches of one step, ncu launch list "
                                         "of round " + str((ncu_traffic() or {}).get("round")),
                         "peak_source": pk["src"],
                         "kernel": "igemm_kernel + wgrad_kernel (tcgen05 conv fwd/dgrad/wgrad)",
                         "algorithmic_gflop_per_image": tot_f / 1e9, "conv_ms_per_step": conv_ms,
                         "conv_share_of_step": conv_ms / step_ms,
                         "per_call_ms": {k: round(v, 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1])[:8]}},
        }
    # secondary workload: RektNet 80x80 bs256 fwd+bwd (second half of the BASELINE metric)
    if not args.no_secondary:
        sec = bench_rektnet(dev, rank, world, steps, warmup, timed)
        if rank == 0:
            out["secondary"] = sec
    # tertiary workload (BASELINE config 5): detect -> NMS -> crop -> RektNet inference latency, one GPU only
    if rank == 0 and world == 1 and not args.no_secondary:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import bench_pipeline

            torch.cuda.empty_cache()
            out["pipeline"] = bench_pipeline.run(iters=20)
        except Exception as e:  # never lose the headline line to the extra workload
            out["pipeline"] = {"error": f"{type(e).__name__}: {e}"}
    if rank == 0 and not args.no_cpu_baseline and world == 1:  # reported at N=1 only
        out["cpu_baseline"] = cpu_baseline()
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


def bench_rektnet(dev, rank, world, steps, warmup, timed):
    import cross_ratio_loss
    import keypoint_net
    from b200cv import synth

    B = 256
    torch.manual_seed(17)
    net = keypoint_net.KeypointNet().to(dev).train()
    x, thm, tpts = (t.to(dev) for t in synth.synth_keypoint_batch(B, seed=rank))
    import contextlib

    with contextlib.redirect_stdout(sys.stderr):  # the reference-compatible constructor prints its settings
        loss_fn = cross_ratio_loss.CrossRatioLoss("l2_heatmap", True, 0.055, 0.038)
    params = list(net.parameters())

    def step():
        for p in params:
            p.grad = None
        hm, pts = net(x)
        loss_fn(hm, pts, thm, tpts)[2].backward()

    for _ in range(warmup):
        step()
    ms, _ = timed(step, steps)
    return {"metric": "images/sec fwd+bwd (RektNet KeypointNet 80x80 bs256 per GPU, l2_heatmap + geo loss)",
            "value": world * B * steps / (ms / 1e3), "unit": "img/s", "ms_per_step": ms / steps}


def cpu_baseline():
    """Oracle (port of the reference's CPU PyTorch path) on this box's host cores, bounded sample."""
    from b200cv import cfg_gen
    from oracle import yolo_oracle as YO

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample_b = 8  # ~10-20 s of host work in all: 1 warm-up + 6 timed steps
    d = tempfile.mkdtemp()
    spec = YO.NetSpec(cfg_gen.write_cfg(d, "darknet53", IMG, IMG, CLASSES))
    params, buffers = YO.init_params(spec, seed=0)
    for p in params.values():
        p.requires_grad_(True)
    x, tg = YO.synth_images(sample_b, IMG, IMG), YO.synth_targets(sample_b, TMAX)

    def step():
        for p in params.values():
            p.grad = None
        YO.darknet_forward(spec, params, buffers, x, tg, LOSS_CONSTS)[0].backward()

    step()
    k = 6
    t0 = time.perf_counter()
    for _ in range(k):
        step()
    dt = (time.perf_counter() - t0) / k
    return {"value": sample_b / dt, "unit": "img/s", "cores": cores, "kind": "port",
            "sample": f"Darknet-53 {IMG}x{IMG} C={CLASSES} fwd+bwd at batch {sample_b}, 1 warm-up + {k} timed steps, "
                      f"torch {torch.__version__} fp32, {cores} threads"}


if __name__ == "__main__":
    main()
