#!/usr/bin/env python
"""Benchmark of the hot path, synthetic data.  Headline = BASELINE.json configs[2]: images/sec, forward+backward,
CVC-YOLOv3 Darknet-53 416x416 bs64 per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c3|c4|c2|rektnet] [--precision bf16|fp32]
    python bench.py --impl reference [--config ...]          # the unmodified reference on the host cores
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

  c3       Darknet-53 416x416 C=80 bs64/GPU                       (default, the headline)
  c4       Darknet-53 608x608 C=80 bs32/GPU                       (BASELINE configs[3])
  c2       YOLOv3-tiny 416x416 C=80 bs16/GPU                      (BASELINE configs[1])
  rektnet  KeypointNet 80x80 bs256/GPU, l2_heatmap + geo loss     (second half of the BASELINE metric)

One JSON line on rank 0.  `value` = whole-job img/s with inputs resident in HBM; `e2e` = the same step through the
public API with pinned-host inputs (H2D inside the timed region, loss read back); `roofline` = the tcgen05 conv
kernels' achieved TFLOP/s (algorithmic conv FLOPs of a step / CUDA-event time of the conv launches) against the
measured sustained bf16 peak -- `frac` over the conv launches alone, `frac_of_step` over the whole step;
`cpu_baseline` = the reference's own modules (baseline/_ref, staged by __graft_entry__.build(); else the oracle
port) timed on this box's host cores.  The default (c3) line also carries the other configurations under `configs`
and the detect -> NMS -> crop -> RektNet latency (BASELINE configs[4]) under `pipeline`.
"""
import argparse
import contextlib
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "mit-driverless-cv-traininginfra_b200")
REF = os.path.join(ROOT, "baseline", "_ref")

import torch  # noqa: E402

TMAX = 16
LOSS_CONSTS = (2.0, 1.6, 25.0, 0.1)  # xy, wh, no_object, object (CVC-YOLOv3/train.py:312-315)
CONFIGS = {
    "c3": dict(kind="yolo", net="darknet53", img=416, batch=64, classes=80, cpu_batch=64,
               metric="images/sec fwd+bwd (YOLOv3 Darknet-53 416x416 bs64 per GPU)"),
    "c4": dict(kind="yolo", net="darknet53", img=608, batch=32, classes=80, cpu_batch=4,
               metric="images/sec fwd+bwd (YOLOv3 Darknet-53 608x608 bs32 per GPU)"),
    "c2": dict(kind="yolo", net="tiny", img=416, batch=16, classes=80, cpu_batch=16,
               metric="images/sec fwd+bwd (YOLOv3-tiny 416x416 bs16 per GPU)"),
    "rektnet": dict(kind="rektnet", batch=256, cpu_batch=32,
                    metric="images/sec fwd+bwd (RektNet KeypointNet 80x80 bs256 per GPU, l2_heatmap + geo loss)"),
}
# sums of per-layer rooflines max(compute, memory) with bf16 activations and the measured peaks
PER_LAYER_ROOFLINE_IMG_S = {"c3": 6000.0, "c4": 2800.0, "rektnet": 85000.0}


def product_paths():
    for p in (ROOT, PKG, os.path.join(PKG, "CVC-YOLOv3"), os.path.join(PKG, "RektNet")):
        if p not in sys.path:
            sys.path.insert(0, p)


def workload(name):
    c = CONFIGS[name]
    if c["kind"] == "yolo":
        label = "Darknet-53" if c["net"] == "darknet53" else "YOLOv3-tiny"
        return f"CVC-YOLOv3 {label} {c['img']}x{c['img']} bs{c['batch']}/GPU fwd+bwd, classes={c['classes']}, T={TMAX}"
    return f"RektNet KeypointNet 80x80 bs{c['batch']}/GPU fwd+bwd, l2_heatmap + geo loss (gamma 0.055/0.038)"


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
            ho = (h + 2 * L["pad"] - L.get("dil", 1) * (L["k"] - 1) - 1) // L["stride"] + 1
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


def rektnet_flops_per_image(size=80):
    """RektNet/keypoint_net.py:17-31: stem 7x7 3->16, four blocks (3x3 dilated, 3x3, 1x1 shortcut), head 1x1 128->7;
    every layer at full resolution.  (forward, fwd+dgrad+wgrad); the head is counted once."""
    px = size * size
    convs = [(3, 16, 7)]
    for cin, cout in ((16, 16), (16, 32), (32, 64), (64, 128)):
        convs += [(cin, cout, 3), (cout, cout, 3), (cin, cout, 1)]
    convs.append((128, 7, 1))
    fwd = sum(2.0 * ci * co * k * k * px for ci, co, k in convs)
    return fwd, 3 * fwd - 2.0 * 3 * 16 * 49 * px


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
def _reference_step_fn(name, batch):
    """(step function, kind, description): one forward+backward of configuration `name` at `batch` images on the CPU
    through the reference's OWN modules (baseline/_ref: models.Darknet / KeypointNet + CrossRatioLoss, unmodified), or
    -- when that copy was never staged -- through the oracle port."""
    c = CONFIGS[name]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sys.path.insert(0, PKG)  # b200cv.cfg_gen / b200cv.synth only (pure host code; no kernels are loaded)
    from b200cv import cfg_gen, synth

    staged = os.path.isfile(os.path.join(REF, "CVC-YOLOv3", "models.py"))
    d = tempfile.mkdtemp()
    if c["kind"] == "yolo":
        cfg = cfg_gen.write_cfg(d, c["net"], c["img"], c["img"], c["classes"])
        x, tg = synth.synth_images(batch, c["img"], c["img"]), synth.synth_targets(batch, TMAX)
        if staged:
            sys.path.insert(0, os.path.join(REF, "CVC-YOLOv3"))
            import models as ref_models  # the reference's file
            from utils.utils import weights_init_normal

            assert os.path.realpath(ref_models.__file__).startswith(os.path.realpath(REF))
            torch.manual_seed(0)
            with contextlib.redirect_stdout(sys.stderr):
                model = ref_models.Darknet(cfg, *LOSS_CONSTS, True)
            model.apply(weights_init_normal)
            model.train()

            def step():
                model.zero_grad()
                losses = model(x, tg)
                losses[0].sum().backward()  # train.py:68-70
            return step, "reference", "reference models.Darknet (baseline/_ref, unmodified)"
        sys.path.insert(0, ROOT)
        from oracle import yolo_oracle as YO

        spec = YO.NetSpec(cfg)
        params, buffers = YO.init_params(spec, seed=0)
        for p in params.values():
            p.requires_grad_(True)

        def step():
            for p in params.values():
                p.grad = None
            YO.darknet_forward(spec, params, buffers, x, tg, LOSS_CONSTS)[0].backward()
        return step, "port", "oracle port of models.Darknet"
    x, thm, tpts = synth.synth_keypoint_batch(batch, seed=0)
    if staged:
        sys.path.insert(0, os.path.join(REF, "RektNet"))
        import cross_ratio_loss as ref_loss
        import keypoint_net as ref_net

        assert os.path.realpath(ref_net.__file__).startswith(os.path.realpath(REF))
        torch.manual_seed(17)
        with contextlib.redirect_stdout(sys.stderr):
            net = ref_net.KeypointNet()
            loss_fn = ref_loss.CrossRatioLoss("l2_heatmap", True, 0.055, 0.038)
        net.train()

        def step():
            net.zero_grad()
            hm, pts = net(x)
            loss_fn(hm, pts, thm, tpts)[2].backward()  # train_eval.py:69-71
        return step, "reference", "reference KeypointNet + CrossRatioLoss (baseline/_ref, unmodified)"
    sys.path.insert(0, ROOT)
    from oracle import rektnet_oracle as RO

    params, buffers = RO.init_params(seed=17)
    for p in params.values():
        p.requires_grad_(True)

    def step():
        for p in params.values():
            p.grad = None
        hm, pts = RO.keypointnet_forward(params, buffers, x, True)
        RO.cross_ratio_loss(hm, pts, thm, tpts, "l2_heatmap", True, 0.055, 0.038)[2].backward()
    return step, "port", "oracle port of KeypointNet + CrossRatioLoss"


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path on this box's host cores, same metric and
    configuration.  A Darknet-53 bs64 step takes ~9 s on 16 cores, so the run is bounded to <= 1 warm-up + <= 2 timed
    steps of the REAL batch (the headline) -- `--cpu-batch` selects a smaller sample for the slower configurations."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    c = CONFIGS[args.config]
    batch = args.cpu_batch or c["cpu_batch"]
    step, kind, what = _reference_step_fn(args.config, batch)
    cores = os.cpu_count() or 1
    for _ in range(max(1, min(args.warmup, 1))):
        step()
    k = max(1, min(args.steps, 2))
    t0 = time.perf_counter()
    for _ in range(k):
        step()
    dt = (time.perf_counter() - t0) / k
    v = batch / dt
    same = batch == c["batch"]
    sample = (f"{what}: fwd+bwd at batch {batch}" + ("" if same else f" (bounded sample of the bs{c['batch']} step)") +
              f", 1 warm-up + {k} timed steps, torch {torch.__version__} fp32, {cores} threads")
    print(json.dumps({
        "impl": "reference", "metric": c["metric"], "value": v, "unit": "img/s", "n_gpus": args.gpus, "steps": k,
        "warmup": 1, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload(args.config), "batch_timed": batch, "same_batch_as_b200_arm": same},
        "cpu_baseline": {"value": v, "unit": "img/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def cpu_baseline(name, cpu_batch=None):
    """The reference arm in its own interpreter (module names collide with the product's), bounded sample."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--config", name, "--steps", "1",
           "--warmup", "1"] + (["--cpu-batch", str(cpu_batch)] if cpu_batch else [])
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    env["CUDA_VISIBLE_DEVICES"] = ""
    try:
        p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
        return json.loads(p.stdout.strip().splitlines()[-1])["cpu_baseline"]
    except Exception as e:
        return {"error": f"{type(e).__name__}: {e}"}


# ----------------------------------------------------------------------------------------------- B200 arm
CONV_ENTRY_POINTS = ("b200cv_conv_fwd", "b200cv_conv_dgrad", "b200cv_conv_dgrad_d2s", "b200cv_conv_wgrad",
                     "b200cv_conv_image_fwd", "b200cv_conv_image_wgrad")


def serial_profile(step_fn):
    """CUDA-event time of every ABI call of ONE eager step with everything on one stream: per-launch events need the
    eager launches (not the captured graph), and a kernel's own duration needs it to run alone (the product path
    queues weight gradients on a side stream, which would be billed to whatever shares the SMs with them)."""
    from b200cv.lib import lib

    saved = {k: os.environ.get(k) for k in ("B200CV_CUDA_GRAPH", "B200CV_WGRAD_TAIL_FILL")}
    os.environ["B200CV_CUDA_GRAPH"] = "0"
    os.environ["B200CV_WGRAD_TAIL_FILL"] = "0"
    try:
        step_fn()  # the first eager call after graph replays may allocate
        return lib().profile_step(step_fn)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


class Harness:
    def __init__(self, dev, rank, world, steps, warmup):
        self.dev, self.rank, self.world, self.steps, self.warmup = dev, rank, world, steps, warmup

    def barrier(self):
        if self.world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, n):
        from b200cv.lib import lib

        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = lib().launches
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        self.barrier()
        ms = e0.elapsed_time(e1)
        if self.world > 1:
            t = torch.tensor([ms], device=self.dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t)
        return ms, lib().launches - l0

    def _result_landing(self, result, slot):
        """Pinned host buffer the step result is copied into (two, alternating)."""
        key = (tuple(result.shape), result.dtype, slot)
        cache = self.__dict__.setdefault("_landing", {})
        if key not in cache:
            cache[key] = torch.empty(result.shape, dtype=result.dtype).pin_memory()
        return cache[key]

    def e2e(self, host_tensors, step_fn, n):
        """Every step copies ITS inputs from pinned host memory (double-buffered on a copy stream, so the copy of step
        i+1 overlaps the compute of step i) and reads the step's result back."""
        bufs = [[torch.empty(t.shape, dtype=t.dtype, device=self.dev) for t in host_tensors] for _ in range(2)]
        copy_stream = torch.cuda.Stream()
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        freed = [torch.cuda.Event(), torch.cuda.Event()]

        def prefetch(slot):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(freed[slot])
                for dst, src in zip(bufs[slot], host_tensors):
                    dst.copy_(src, non_blocking=True)
                ready[slot].record(copy_stream)

        def run(count):
            for ev in freed:
                ev.record()
            prefetch(0)
            host, pending = None, None
            for i in range(count):
                slot = i & 1
                if i + 1 < count:
                    prefetch(slot ^ 1)
                torch.cuda.current_stream().wait_event(ready[slot])
                result = step_fn(*bufs[slot])
                freed[slot].record()
                # device -> host read of EVERY step's result: the copy is queued behind the step and consumed one
                # step later (what a training loop that logs the previous step's loss does), so the host enqueues
                # step i+1 while step i runs; the last result is read inside the timed region too
                landing = self._result_landing(result, i & 1)
                landing.copy_(result, non_blocking=True)
                done = torch.cuda.Event()
                done.record()
                if pending is not None:
                    pending[1].synchronize()
                    host = pending[0].clone()
                pending = (landing, done)
            pending[1].synchronize()
            return pending[0].clone()

        run(2)
        ms, _ = self.timed(lambda: run(n), 1)
        return ms, int(sum(t.numel() * t.element_size() for t in host_tensors))


def bench_yolo(name, H, precision, with_clocks=True, ncu_window=False):
    from b200cv import cfg_gen, synth
    from b200cv.lib import lib
    import models
    from utils.utils import weights_init_normal

    c = CONFIGS[name]
    IMG, BATCH = c["img"], c["batch"]
    d = tempfile.mkdtemp()
    torch.manual_seed(0)
    model = models.Darknet(cfg_gen.write_cfg(d, c["net"], IMG, IMG, c["classes"]), *LOSS_CONSTS, True)
    model.apply(weights_init_normal)
    model = model.to(H.dev).train()
    model.engine().set_precision(precision)
    params = list(model.parameters())
    imgs_h = synth.synth_images(BATCH, IMG, IMG, seed=H.rank).pin_memory()
    tg_h = synth.synth_targets(BATCH, TMAX, seed=1 + H.rank).pin_memory()
    imgs_d, tg_d = imgs_h.to(H.dev), tg_h.to(H.dev)

    def step(x, t):
        for p in params:
            p.grad = None
        losses = model(x, t)
        losses[0].sum().backward()
        return losses

    if ncu_window:
        os.environ["B200CV_CUDA_GRAPH"] = "0"
    for _ in range(H.warmup):
        step(imgs_d, tg_d)
    if ncu_window:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step(imgs_d, tg_d)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return None
    sampler = ClockSampler(H.dev.index) if (H.rank == 0 and with_clocks) else None
    ms, launches = H.timed(lambda: step(imgs_d, tg_d), H.steps)
    clocks = sampler.stop() if sampler else None
    value = H.world * BATCH * H.steps / (ms / 1e3)
    ms_e2e, h2d = H.e2e([imgs_h, tg_h], lambda x, t: torch.stack([l.detach() for l in step(x, t)]), H.steps)
    e2e_value = H.world * BATCH * H.steps / (ms_e2e / 1e3)
    # roofline of the dominant kernels: event-time every conv launch of one more step (every rank runs the step -- it
    # contains the gradient all-reduce -- rank 0 reports)
    prof = serial_profile(lambda: step(imgs_d, tg_d))
    out = None
    if H.rank == 0:
        fwd_f, tot_f = conv_flops_per_image(synth.conv_layer_table(model), IMG)
        pk = peaks()
        conv_ms = sum(v for k, v in prof.items() if k in CONV_ENTRY_POINTS)
        step_ms = sum(prof.values())
        achieved = tot_f * BATCH / (conv_ms / 1e3) / 1e12
        of_step = tot_f * BATCH / (ms / H.steps / 1e3) / 1e12
        traffic = ncu_traffic() if name == "c3" and precision == "bf16" else None
        roof = {"bound": "tensor", "achieved": achieved, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                "frac": achieved / pk["bf16_sustained"],
                "frac_of_step": of_step / pk["bf16_sustained"], "achieved_over_step": of_step,
                "traffic": (traffic or {}).get("conv_dram_bytes_per_step"),
                "traffic_note": "DRAM bytes (read+write) of all conv launches of one step, ncu launch list of round "
                                + str((traffic or {}).get("round")),
                "peak_source": pk["src"], "kernel": "igemm_kernel + wgrad_kernel (tcgen05 conv fwd/dgrad/wgrad)",
                "algorithmic_gflop_per_image": tot_f / 1e9, "conv_ms_per_step": conv_ms,
                "conv_share_of_step": conv_ms / step_ms,
                "per_call_ms": {k: round(v, 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1])[:8]}}
        if name in PER_LAYER_ROOFLINE_IMG_S:
            roof["frac_of_per_layer_roofline"] = value / H.world / PER_LAYER_ROOFLINE_IMG_S[name]
        out = {"metric": c["metric"], "value": value, "unit": "img/s", "n_gpus": H.world, "steps": H.steps,
               "warmup": H.warmup, "ms_per_step": ms / H.steps, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None,
               "dtype": "bf16" if precision == "bf16" else "bf16x6 (fp32-parity mode: three bf16 pieces per value, "
                                                            "six tensor-core passes, fp32 accumulation)",
               "data": "synthetic",
               "config": {"workload": workload(name), "parallelism": f"dp{H.world}", "global_batch": BATCH * H.world,
                          "l2": "activations per step (GBs) exceed the 126 MB L2; no explicit flush"},
               "e2e": {"value": e2e_value, "unit": "img/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 28,
                       "ms_per_step": ms_e2e / H.steps},
               "gpu_launches": launches, "clocks": clocks, "roofline": roof}
    del model
    torch.cuda.empty_cache()
    return out


def bench_rektnet(H, precision, with_clocks=False):
    import cross_ratio_loss
    import keypoint_net
    from b200cv import synth
    from b200cv.lib import lib

    c = CONFIGS["rektnet"]
    B = c["batch"]
    torch.manual_seed(17)
    net = keypoint_net.KeypointNet().to(H.dev).train()
    net.engine().set_precision(precision)
    x_h, thm_h, tpts_h = (t.pin_memory() for t in synth.synth_keypoint_batch(B, seed=H.rank))
    x, thm, tpts = x_h.to(H.dev), thm_h.to(H.dev), tpts_h.to(H.dev)
    with contextlib.redirect_stdout(sys.stderr):  # the reference-compatible constructor prints its settings
        loss_fn = cross_ratio_loss.CrossRatioLoss("l2_heatmap", True, 0.055, 0.038)
    params = list(net.parameters())

    def step(x_, thm_, tpts_):
        for p in params:
            p.grad = None
        hm, pts = net(x_)
        loss3 = loss_fn(hm, pts, thm_, tpts_)
        loss3[2].backward()
        return loss3[2].detach()

    for _ in range(H.warmup):
        step(x, thm, tpts)
    sampler = ClockSampler(H.dev.index) if (H.rank == 0 and with_clocks) else None
    ms, launches = H.timed(lambda: step(x, thm, tpts), H.steps)
    clocks = sampler.stop() if sampler else None
    value = H.world * B * H.steps / (ms / 1e3)
    ms_e2e, h2d = H.e2e([x_h, thm_h, tpts_h], lambda a, b, c_: step(a, b, c_).view(1), H.steps)
    prof = serial_profile(lambda: step(x, thm, tpts))
    out = None
    if H.rank == 0:
        fwd_f, tot_f = rektnet_flops_per_image()
        pk = peaks()
        conv_ms = sum(v for k, v in prof.items() if k in CONV_ENTRY_POINTS)
        achieved = tot_f * B / (conv_ms / 1e3) / 1e12
        of_step = tot_f * B / (ms / H.steps / 1e3) / 1e12
        out = {"metric": c["metric"], "value": value, "unit": "img/s", "n_gpus": H.world, "steps": H.steps,
               "warmup": H.warmup, "ms_per_step": ms / H.steps, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "bf16" if precision == "bf16" else "bf16x6 (fp32-parity mode)",
               "data": "synthetic",
               "config": {"workload": workload("rektnet"), "parallelism": f"dp{H.world}", "global_batch": B * H.world,
                          "l2": "activations per step (GBs) exceed the 126 MB L2; no explicit flush"},
               "e2e": {"value": H.world * B * H.steps / (ms_e2e / 1e3), "unit": "img/s", "h2d_bytes_per_step": h2d,
                       "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / H.steps},
               "gpu_launches": launches, "clocks": clocks,
               "roofline": {"bound": "tensor", "achieved": achieved, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                            "frac": achieved / pk["bf16_sustained"], "frac_of_step": of_step / pk["bf16_sustained"],
                            "frac_of_per_layer_roofline": value / H.world / PER_LAYER_ROOFLINE_IMG_S["rektnet"],
                            "traffic": None, "peak_source": pk["src"],
                            "note": "11 of the 14 convolutions (16/32-channel layers at 80x80) are HBM-bound in bf16 "
                                    "(SURVEY 8a-12): the per-layer roofline, not the tensor peak, is the ceiling",
                            "kernel": "igemm_kernel + wgrad_kernel", "algorithmic_gflop_per_image": tot_f / 1e9,
                            "conv_ms_per_step": conv_ms,
                            "per_call_ms": {k: round(v, 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1])[:8]}}}
    del net
    torch.cuda.empty_cache()
    return out


def bench_pipeline_all_ranks(H):
    """BASELINE configs[4]: detect -> NMS -> crop -> RektNet at batch 128, every rank an independent replica (the path
    has no exchange step); rank 0 reports its own line plus the spread over the replicas."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_pipeline

    try:
        torch.cuda.empty_cache()
        line = bench_pipeline.run(iters=20)
    except Exception as e:  # never lose the headline line to the extra workload
        line = {"error": f"{type(e).__name__}: {e}"}
    if H.world > 1:
        lines = [None] * H.world
        torch.distributed.all_gather_object(lines, line)
        if H.rank == 0 and "error" not in line:
            ok = [ln for ln in lines if ln and "error" not in ln]
            line["n_gpus"] = H.world
            line["replicas"] = {
                "p50_ms_max": max(ln["resident"]["p50_ms"] for ln in ok),
                "p99_ms_max": max(ln["resident"]["p99_ms"] for ln in ok),
                "from_frames_e2e_p50_ms_max": max(ln["from_frames"]["e2e"]["p50_ms"] for ln in ok),
                "from_frames_e2e_p99_ms_max": max(ln["from_frames"]["e2e"]["p99_ms"] for ln in ok),
                "images_per_s_sum_at_p50": sum(ln["resident"]["images_per_s_at_p50"] for ln in ok),
                "n_ok": len(ok)}
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS))
    ap.add_argument("--precision", default=os.environ.get("B200CV_PRECISION", "bf16"), choices=("bf16", "fp32"))
    ap.add_argument("--cpu-batch", type=int, default=0, help="batch of the CPU reference sample (0 = per-config default)")
    ap.add_argument("--only", action="store_true", help="just the selected configuration: no other configs / pipeline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="alias of --only")
    ap.add_argument("--ncu-window", action="store_true",
                    help="after warm-up run ONE eager step between cudaProfilerStart/Stop and exit "
                         "(for `ncu --profile-from-start off`; prints no bench line)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    product_paths()
    from b200cv import parallel

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback")
    local = parallel.init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    H = Harness(dev, parallel.rank(), parallel.world_size(), max(1, args.steps), max(3, args.warmup))
    only = args.only or args.no_secondary
    name = args.config
    if CONFIGS[name]["kind"] == "yolo":
        out = bench_yolo(name, H, args.precision, ncu_window=args.ncu_window)
        if args.ncu_window:
            return
    else:
        out = bench_rektnet(H, args.precision, with_clocks=True)
    if not only and name == "c3" and args.precision == "bf16":
        # the other named configurations, each a full line of its own (shorter runs), and the inference joint
        others = {}
        Hs = Harness(dev, H.rank, H.world, max(3, min(H.steps, 10)), 3)
        for other in ("c4", "c2"):
            others[other] = bench_yolo(other, Hs, "bf16", with_clocks=False)
        others["rektnet"] = bench_rektnet(Hs, "bf16")
        others["c3_fp32_parity_mode"] = bench_yolo("c3", Harness(dev, H.rank, H.world, 3, 3), "fp32", with_clocks=False)
        pipe = bench_pipeline_all_ranks(H)
        if H.rank == 0:
            out["configs"] = others
            out["secondary"] = others["rektnet"]
            out["pipeline"] = pipe
    if H.rank == 0 and not args.no_cpu_baseline and H.world == 1:  # reported at N=1 only
        out["cpu_baseline"] = cpu_baseline(name, args.cpu_batch or None)
        if "configs" in out:
            for other in ("c4", "c2", "rektnet"):
                if out["configs"].get(other):
                    out["configs"][other]["cpu_baseline"] = cpu_baseline(other)
    if H.rank == 0:
        print(json.dumps(out))
    if H.world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
