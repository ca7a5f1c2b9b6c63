"""The fp32-parity mode (B200CV_PRECISION=fp32 / engine.set_precision("fp32")): operands split into three bf16
pieces (24 mantissa bits), six tensor-core passes, fp32 accumulation.  The reference path is fp32 end to end (CVC-YOLOv3/models.py:59-69, TF32 never enabled) and the north star
states 1e-3 relative for it; these tests hold the B200 path to that bar against the REFERENCE goldens -- per kernel
(teacher-forced, against torch fp32 on the same operands) and for whole networks including every parameter gradient."""
import pytest
import torch
import torch.nn.functional as F

import helpers
from b200cv import ops
from oracle import rektnet_oracle as RO
from oracle import yolo_oracle as YO

pytestmark = pytest.mark.gpu
DEV = "cuda"
FP32_RTOL = 1e-3  # BASELINE.json north star: "conv activations and losses within 1e-3 rel fp32"


def _rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


CASES = [  # N, H, W, Cin, Cout, k, stride, pad, dil
    (2, 13, 13, 64, 128, 3, 1, 1, 1),
    (3, 26, 26, 128, 256, 3, 1, 1, 1),
    (2, 16, 16, 32, 64, 3, 2, 1, 1),
    (2, 13, 13, 64, 64, 3, 2, 1, 1),
    (2, 20, 20, 16, 16, 3, 1, 2, 2),
    (2, 13, 13, 256, 18, 1, 1, 0, 1),
    (2, 13, 13, 1024, 255, 1, 1, 0, 1),
    (2, 32, 32, 3, 32, 3, 1, 1, 1),
    (4, 13, 13, 512, 1024, 3, 1, 1, 1),
    (2, 19, 19, 128, 64, 1, 1, 0, 1),
]


@pytest.mark.parametrize("case", CASES)
def test_split_conv_fwd_dgrad_wgrad_vs_torch_fp32(case):
    """Teacher-forced per-kernel parity: fp32 operands (NOT pre-rounded to bf16), every result within 1e-5 of torch
    fp32 (norm-wise) -- three orders below the bf16 mode; split outputs carry 24 mantissa bits."""
    n, h, w, cin, cout, k, s, p, d = case
    g = torch.Generator().manual_seed(sum(case))
    x = torch.randn(n, cin, h, w, generator=g).requires_grad_(True)
    wt = (torch.randn(cout, cin, k, k, generator=g) * 0.1).requires_grad_(True)
    y_ref = F.conv2d(x, wt, None, s, p, d)
    dy = torch.randn(y_ref.shape, generator=g)
    y_ref.backward(dy)
    with ops.precision(True):
        xd = ops.nchw_to_nhwc(x.detach().to(DEV))
        assert xd.shape[-1] == ops.split_pieces() * ops.pad_channels(cin)
        wpk = ops.pack_weights(wt.detach().to(DEV), False)
        wpk_t = ops.pack_weights(wt.detach().to(DEV), True)
        y = ops.conv_fwd(xd, wpk, cout, k, s, p, d)
        assert _rel(ops.nhwc_to_nchw(y, cout).cpu(), y_ref.detach()) < 1e-5
        y32 = ops.conv_fwd(xd, wpk, cout, k, s, p, d, out_dtype=torch.float32)  # fp32 head output
        assert _rel(y32[..., :cout].permute(0, 3, 1, 2).cpu(), y_ref.detach()) < 1e-5
        dyd = ops.nchw_to_nhwc(dy.to(DEV))
        dx = ops.conv_dgrad(dyd, wpk_t, cin, k, s, p, d, (h, w))
        assert _rel(ops.nhwc_to_nchw(dx, cin).cpu(), x.grad) < 1e-5
        res = ops.nchw_to_nhwc(torch.randn(n, cin, h, w, generator=g).to(DEV))
        want = ops.nhwc_to_nchw(dx, cin) + ops.nhwc_to_nchw(res, cin)
        dx2 = ops.conv_dgrad(dyd, wpk_t, cin, k, s, p, d, (h, w), out=res, residual=res)  # gradient fan-in
        assert _rel(ops.nhwc_to_nchw(dx2, cin), want) < 1e-5
        dwp = ops.conv_wgrad(xd, dyd, cout, k, s, p, d)
        dw = torch.empty(cout, cin, k, k, device=DEV)
        ops.unpack_wgrad(dwp, dw)
        assert _rel(dw.cpu(), wt.grad) < 1e-5


def test_split_epilogue_statistics_affine_residual():
    n, h, w, cin, cout = 4, 13, 13, 64, 128
    g = torch.Generator().manual_seed(7)
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, 3, 3, generator=g) * 0.1
    ref = F.conv2d(x, wt, None, 1, 1).permute(0, 2, 3, 1)
    with ops.precision(True):
        xd, wpk = ops.nchw_to_nhwc(x.to(DEV)), ops.pack_weights(wt.to(DEV), False)
        stats = ops.stats_buffer(cout, DEV)
        y = ops.split_to_f32(ops.conv_fwd(xd, wpk, cout, 3, 1, 1, stats=stats))
        tot = ops.stats_value(stats)
        assert torch.allclose(tot[:cout].float().cpu(), ref.sum((0, 1, 2)), rtol=1e-4, atol=1e-3)
        assert torch.allclose(tot[cout:].float().cpu(), (ref * ref).sum((0, 1, 2)), rtol=1e-4, atol=1e-3)
        assert _rel(y.cpu(), ref) < 1e-4
        scale = torch.rand(cout, device=DEV) + 0.5
        shift = torch.randn(cout, device=DEV)
        res32 = torch.randn(n, h, w, cout, device=DEV)
        res = ops.split_from_f32(res32)
        assert _rel(ops.split_to_f32(res), res32) < 1e-5
        for after in (False, True):
            out = ops.split_to_f32(ops.conv_fwd(xd, wpk, cout, 3, 1, 1, scale=scale, shift=shift, residual=res,
                                                act=ops.ACT_LEAKY, slope=0.1, res_after_act=after))
            z = ref.to(DEV) * scale + shift
            want = F.leaky_relu(z, 0.1) + res32 if after else F.leaky_relu(z + res32, 0.1)
            assert _rel(out, want) < 1e-4


def test_split_bn_pool_upsample_against_torch():
    n, h, w, c = 3, 12, 12, 64
    g = torch.Generator().manual_seed(11)
    y = torch.randn(n, c, h, w, generator=g) * 2 + 0.3
    gamma, beta = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g)
    da = torch.randn(n, c, h, w, generator=g)
    yr = y.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    a_ref = F.leaky_relu(F.batch_norm(yr, torch.zeros(c), torch.ones(c), gr, br, True, 0.1, 1e-5), 0.1)
    a_ref.backward(da)
    f = lambda k: torch.empty(k, device=DEV)
    with ops.precision(True):
        yd = ops.nchw_to_nhwc(y.to(DEV))
        y32 = ops.split_to_f32(yd)
        stats = ops.stats_encode(torch.stack([y32.sum((0, 1, 2)), (y32 ** 2).sum((0, 1, 2))]).flatten())
        scale, shift, mean, rstd, coef = f(c), f(c), f(c), f(c), f(3 * c)
        rmd, rvd = torch.zeros(c, device=DEV), torch.ones(c, device=DEV)
        a = ops.bn_stats_apply_act(stats, n * h * w, gamma.to(DEV), beta.to(DEV), None, 1e-5, 0.1, rmd, rvd, scale,
                                   shift, mean, rstd, yd, ops.ACT_LEAKY, 0.1)
        assert _rel(ops.nhwc_to_nchw(a, c).cpu(), a_ref.detach()) < 1e-4
        dad = ops.nchw_to_nhwc(da.to(DEV))
        sums = ops.bn_bwd_reduce(dad, yd, None, scale, shift, mean, rstd, ops.ACT_LEAKY, 0.1)
        dg, db = f(c), f(c)
        dy = ops.bn_bwd_stats_apply(sums, n * h * w, gamma.to(DEV), coef, dg, db, dad, yd, scale, shift, mean, rstd,
                                    ops.ACT_LEAKY, 0.1)
        assert torch.allclose(dg.cpu(), gr.grad, rtol=1e-3, atol=1e-3)
        assert torch.allclose(db.cpu(), br.grad, rtol=1e-3, atol=1e-3)
        assert _rel(ops.nhwc_to_nchw(dy, c).cpu(), yr.grad) < 2e-4
        # the same reduction with the activation sign taken from the saved output (RektNet blocks)
        sums2 = ops.bn_bwd_reduce(dad, yd, a, None, None, mean, rstd, ops.ACT_LEAKY, 0.1)
        assert torch.allclose(ops.stats_value(sums2), ops.stats_value(sums), rtol=1e-6, atol=1e-6)
        assert _rel(ops.act_bwd(dad, a, ops.ACT_LEAKY, 0.1).float(), (ops.split_from_f32(
            ops.split_to_f32(dad) * torch.where(ops.split_to_f32(a) > 0, 1.0, 0.1))).float()) < 1e-6

        xp = torch.randn(n, c, h, w, generator=g).requires_grad_(True)
        xd = ops.nchw_to_nhwc(xp.detach().to(DEV))
        x16 = ops.nhwc_to_nchw(xd, c).cpu()  # what the split tensor holds (16-bit mantissa)
        for stride in (2, 1):
            xq = x16.clone().requires_grad_(True)
            ref = F.max_pool2d(F.pad(xq, (0, 1, 0, 1)) if stride == 1 else xq, 2, stride)
            gy = torch.randn(ref.shape, generator=g)
            ref.backward(gy)
            out = ops.maxpool_fwd(xd, stride)
            assert torch.equal(ops.nhwc_to_nchw(out, c).cpu(), ref.detach())
            dx = ops.maxpool_bwd(xd, ops.nchw_to_nhwc(gy.to(DEV)), stride)
            assert _rel(ops.nhwc_to_nchw(dx, c).cpu(), xq.grad) < 1e-4
        up = ops.upsample_fwd(xd)
        assert torch.equal(ops.nhwc_to_nchw(up, c).cpu(), F.interpolate(x16, scale_factor=2, mode="nearest"))
        gu = torch.randn(n, c, 2 * h, 2 * w, generator=g)
        dxu = ops.upsample_bwd(ops.nchw_to_nhwc(gu.to(DEV)))
        assert _rel(ops.nhwc_to_nchw(dxu, c).cpu(), gu.view(n, c, h, 2, w, 2).sum((3, 5))) < 1e-4
        # concat / slice helpers (route layers)
        a32, b32 = torch.randn(n, h, w, 64, device=DEV), torch.randn(n, h, w, 128, device=DEV)
        cat = ops.concat_channels([ops.split_from_f32(a32), ops.split_from_f32(b32)])
        assert cat.shape[-1] == ops.split_pieces() * 192
        assert _rel(ops.split_to_f32(cat), torch.cat([a32, b32], -1)) < 1e-5
        assert _rel(ops.split_to_f32(ops.slice_grad(cat, 64, 128)), b32) < 1e-5
        acc = ops.split_from_f32(a32)
        ops.slice_grad(cat, 0, 64, into=acc)
        assert _rel(ops.split_to_f32(acc), 2 * a32) < 1e-5
        assert torch.allclose(ops.bias_grad(ops.split_from_f32(b32), 100), b32.sum((0, 1, 2))[:100], rtol=1e-4,
                              atol=1e-3)


def _run_fp32(cfg_dir, g):
    model, path = helpers.make_darknet(cfg_dir, g["cfg"], g["S"], g["C"])
    model = model.to(DEV).train()
    model.engine().set_precision("fp32")
    x = YO.synth_images(g["B"], g["S"], g["S"], seed=0).to(DEV)
    tg = YO.synth_targets(g["B"], 16, seed=1).to(DEV)
    losses = model(x, tg)
    losses[0].sum().backward()
    return model, x, tg, losses


@pytest.mark.parametrize("name", ["tiny_128", "tiny_416", "full_128", "tiny_128_c80"])
def test_darknet_fp32_mode_vs_reference_goldens(cfg_dir, golden_yolo, name):
    """Whole-network parity with the REFERENCE goldens in the fp32-parity mode: the 7-tuple (total AND every part)
    within 1e-3 relative (measured: 1e-5), the BN running statistics, the eval-mode detections that follow the step, and
    every parameter gradient by norm and sampled values.  The gradient bounds are what the reference's own fp32
    rounding allows: its CPU fp32 gradients differ from an fp64 evaluation of the same graph by up to 1e-2 on single
    tensors (see test_darknet_fp32_mode_gradients_vs_fp64_yardstick, which measures exactly that)."""
    g = golden_yolo["darknet"][name]
    model, x, tg, losses = _run_fp32(cfg_dir, g)
    got = torch.stack([l.detach() for l in losses]).cpu()
    rel = (got - g["losses"]).abs() / g["losses"].abs().clamp_min(1e-3)
    assert float(rel.max()) < 1e-4 < FP32_RTOL, (got, g["losses"], rel)
    errs = helpers.grad_errors(model.named_parameters(), g["grads"])
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:3]
    vals = sorted(errs.values())
    assert worst[0][1] < 5e-2 and vals[len(vals) // 2] < 1e-2, (worst, vals[len(vals) // 2])
    for k, p in model.named_parameters():
        ref = g["grads"][k]["norm"]
        assert abs(float(p.grad.double().norm()) - ref) <= 2e-2 * ref + 1e-7, (k, float(p.grad.double().norm()), ref)
    for k, (s_, a_) in g["running"].items():
        buf = dict(model.named_buffers())[k]
        assert abs(float(buf.double().sum()) - s_) <= 1e-4 * (a_ + 1), k
    if "det_rows" in g:
        model.eval()
        with torch.no_grad():
            det = model(x)
        got_d, want_d = det.cpu()[:, ::53], g["det_rows"]
        assert float((got_d[..., 4:] - want_d[..., 4:]).abs().max()) < 1e-4
        assert float(((got_d[..., :4] - want_d[..., :4]).abs() / (want_d[..., :4].abs() + 1.0)).max()) < 1e-3


def _oracle_grads(path, model, x, tg, dtype):
    params = {k: v.detach().clone().to(dtype).requires_grad_(True) for k, v in model.named_parameters()}
    buffers = {k: (v.clone().to(dtype) if v.dtype.is_floating_point else v.clone()) for k, v in model.named_buffers()}
    out = YO.darknet_forward(YO.NetSpec(path), params, buffers, x.to(dtype), tg)
    out[0].backward()
    return [float(v) for v in out], {k: v.grad.double() for k, v in params.items()}


@pytest.mark.parametrize("cfg_name,S,B", [("yolo_baseline_tiny.cfg", 416, 4), ("yolo_baseline.cfg", 128, 2)])
def test_darknet_fp32_mode_gradients_vs_fp64_yardstick(cfg_dir, cfg_name, S, B):
    """How close is "fp32 parity" for gradients?  The yardstick is the reference's OWN arithmetic: the oracle
    (restatement pinned to the reference) evaluated in fp32 differs from the same graph evaluated in fp64 by
    e_ref = |g32 - g64| / |g64| per parameter (median 5e-4 / 2.6e-3, worst 1e-2 / 1.7e-2 for these two networks at
    random init: LeakyReLU sign ties).  The B200 fp32-parity mode is held to the same distance from the fp64 ground
    truth: e_b200 within 3x of e_ref (median and worst), every gradient direction cos >= 0.998 (measured: tiny 416
    e_ref 1.2e-3 / 2.3e-3, e_b200 3.9e-4 / 2.2e-3, cos >= 0.999997; Darknet-53 128 -- 4x4 final grid, BatchNorm over 32
    samples, the most ill-conditioned point in the suite -- e_ref 8e-3 / 2.1e-2, e_b200 2e-2 / 4.6e-2, cos >= 0.9989;
    the bf16 mode reaches cos 0.5-0.6 there)."""
    model, path = helpers.make_darknet(cfg_dir, cfg_name, S, 1, seed=2)
    x, tg = YO.synth_images(B, S, S, seed=5), YO.synth_targets(B, 16, seed=6)
    l64, g64 = _oracle_grads(path, model, x, tg, torch.float64)
    l32, g32 = _oracle_grads(path, model, x, tg, torch.float32)
    model = model.to(DEV).train()
    model.engine().set_precision("fp32")
    got = model(x.to(DEV), tg.to(DEV))
    got[0].backward()
    for a, b in zip(got, l64):
        assert abs(float(a) - b) <= 1e-4 * max(abs(b), 1e-3), (float(a), b)
    e_ref, e_b200, cos = [], [], []
    for k, p in model.named_parameters():
        t = g64[k]
        if float(t.norm()) < 1e-12:
            continue
        mine = p.grad.detach().cpu().double()
        e_ref.append(float((g32[k] - t).norm() / t.norm()))
        e_b200.append(float((mine - t).norm() / t.norm()))
        cos.append(float((mine * t).sum() / (mine.norm() * t.norm())))
    med = lambda v: sorted(v)[len(v) // 2]
    summary = dict(ref_median=med(e_ref), ref_max=max(e_ref), b200_median=med(e_b200), b200_max=max(e_b200),
                   cos_min=min(cos))
    print("fp64 yardstick", cfg_name, S, summary)
    assert med(e_b200) <= 3 * med(e_ref) + 1e-5, summary
    assert max(e_b200) <= 3 * max(e_ref) + 1e-4, summary
    assert min(cos) >= 0.998, summary


def test_fp32_mode_is_reproducible_and_graph_replay_matches(cfg_dir):
    model, _ = helpers.make_darknet(cfg_dir, "yolo_baseline_tiny.cfg", 128, 1)
    model = model.to(DEV).train()
    model.engine().set_precision("fp32")
    x, tg = YO.synth_images(2, 128, 128).to(DEV), YO.synth_targets(2, 16).to(DEV)
    opt = torch.optim.SGD(model.parameters(), lr=0.0)
    hist = []
    for _ in range(5):  # two eager steps, then CUDA-graph replays
        opt.zero_grad()
        out = model(x, tg)
        out[0].backward()
        hist.append((torch.stack([o.detach() for o in out]).cpu(),
                     model.module_list[0][0].weight.grad.detach().cpu().clone()))
    for h in hist[1:]:
        assert torch.equal(h[0], hist[0][0])  # integer-accumulated statistics: bit-identical losses
        assert _rel(h[1], hist[0][1]) < 1e-5


@pytest.mark.parametrize("loss_type", ["l2_softargmax", "l2_heatmap", "l1_softargmax"])
@pytest.mark.parametrize("geo", [False, True])
def test_rektnet_fp32_mode_vs_reference(golden_rekt, loss_type, geo):
    """KeypointNet + CrossRatioLoss in the fp32-parity mode vs the reference goldens at 1e-3."""
    import cross_ratio_loss
    import keypoint_net

    g = golden_rekt[f"{loss_type}_geo{int(geo)}"]
    torch.manual_seed(17)
    net = keypoint_net.KeypointNet().to(DEV).train()
    net.engine().set_precision("fp32")
    xc, thmc, tptsc = RO.synth_batch(g["B"], seed=0)
    hm, pts = net(xc.to(DEV))
    loc, geo_l, total = cross_ratio_loss.CrossRatioLoss(loss_type, geo, 0.055, 0.038)(hm, pts, thmc.to(DEV),
                                                                                       tptsc.to(DEV))
    total.backward()
    assert abs(float(loc) - float(g["loc"])) <= 1e-4 * abs(float(g["loc"]))  # north star: 1e-3
    assert abs(float(geo_l) - float(g["geo"])) <= 1e-4 * abs(float(g["geo"])) + 1e-7
    assert abs(float(total) - float(g["total"])) <= 1e-4 * abs(float(g["total"]))
    assert float((pts.detach().cpu() - g["pts"]).abs().max()) < 1e-5
    # analytically zero gradients (conv biases under train-mode BN; the head bias under the softmax): rounding noise
    skip = ("conv.bias", "conv1.bias", "conv2.bias", "shortcut_conv.bias", "out.bias")
    errs = helpers.grad_errors([(k, p) for k, p in net.named_parameters() if not k.endswith(skip)], g["grads"])
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:3]
    vals = sorted(errs.values())
    # the reference's own fp32 gradients sit 2e-3 (norm-wise; 9e-3 on single elements) from an fp64 evaluation
    assert worst[0][1] < 3e-2 and vals[len(vals) // 2] < 5e-3, (worst, vals[len(vals) // 2])
    for k, p in net.named_parameters():
        if k.endswith(skip):
            continue
        ref = g["grads"][k]["norm"]
        assert abs(float(p.grad.double().norm()) - ref) <= 1e-2 * ref + 1e-7, k
