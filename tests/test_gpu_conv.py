"""tcgen05 convolution kernels (through the C ABI) against torch fp32 on the same bf16-rounded operands."""
import pytest
import torch
import torch.nn.functional as F

from b200cv import ops

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _bf(x):
    return x.to(torch.bfloat16).float()


CASES = [  # N, H, W, Cin, Cout, k, stride, pad, dil
    (2, 13, 13, 64, 128, 3, 1, 1, 1),
    (3, 26, 26, 128, 256, 3, 1, 1, 1),
    (2, 16, 16, 32, 64, 3, 2, 1, 1),
    (2, 13, 13, 64, 64, 3, 2, 1, 1),
    (2, 20, 20, 16, 16, 3, 1, 2, 2),
    (2, 20, 20, 3, 16, 7, 1, 3, 1),
    (2, 13, 13, 256, 18, 1, 1, 0, 1),
    (2, 13, 13, 1024, 255, 1, 1, 0, 1),
    (2, 32, 32, 3, 32, 3, 1, 1, 1),
    (4, 13, 13, 512, 1024, 3, 1, 1, 1),
]


@pytest.mark.parametrize("case", CASES)
def test_conv_fwd_dgrad_wgrad(case):
    n, h, w, cin, cout, k, s, p, d = case
    g = torch.Generator().manual_seed(sum(case))
    x = _bf(torch.randn(n, cin, h, w, generator=g)).requires_grad_(True)
    wt = _bf(torch.randn(cout, cin, k, k, generator=g) * 0.1).requires_grad_(True)
    y_ref = F.conv2d(x, wt, None, s, p, d)
    dy = _bf(torch.randn(y_ref.shape, generator=g))
    y_ref.backward(dy)

    xd = ops.nchw_to_nhwc(x.detach().to(DEV))
    wpk = ops.pack_weights(wt.detach().to(DEV), False)
    wpk_t = ops.pack_weights(wt.detach().to(DEV), True)
    y = ops.conv_fwd(xd, wpk, cout, k, s, p, d)
    y_nchw = ops.nhwc_to_nchw(y, cout).cpu()
    tol = 1.5e-2 * float(y_ref.abs().max())
    assert float((y_nchw - y_ref.detach()).abs().max()) <= tol

    dyd = ops.nchw_to_nhwc(dy.to(DEV))
    dx = ops.conv_dgrad(dyd, wpk_t, cin, k, s, p, d, (h, w))
    dx_nchw = ops.nhwc_to_nchw(dx, cin).cpu()
    assert float((dx_nchw - x.grad).abs().max()) <= 1.5e-2 * float(x.grad.abs().max())

    dwp = ops.conv_wgrad(xd, dyd, cout, k, s, p, d)
    dw = torch.empty(cout, cin, k, k, device=DEV)
    ops.unpack_wgrad(dwp, dw)
    assert float((dw.cpu() - wt.grad).abs().max()) <= 2e-3 * float(wt.grad.abs().max())


def test_conv_epilogue_stats_affine_residual():
    n, h, w, cin, cout = 4, 13, 13, 64, 128
    g = torch.Generator().manual_seed(7)
    x = _bf(torch.randn(n, cin, h, w, generator=g))
    wt = _bf(torch.randn(cout, cin, 3, 3, generator=g) * 0.1)
    xd, wpk = ops.nchw_to_nhwc(x.to(DEV)), ops.pack_weights(wt.to(DEV), False)
    stats = ops.stats_buffer(cout, DEV)
    y = ops.conv_fwd(xd, wpk, cout, 3, 1, 1, stats=stats).float()
    tot = ops.stats_value(stats).float()
    assert torch.allclose(tot[:cout], y.sum((0, 1, 2)), rtol=1e-3, atol=1e-2)
    assert torch.allclose(tot[cout:], (y * y).sum((0, 1, 2)), rtol=1e-3, atol=1e-2)
    again = ops.stats_buffer(cout, DEV)
    ops.conv_fwd(xd, wpk, cout, 3, 1, 1, stats=again)
    assert torch.equal(stats.sum(0), again.sum(0))  # integer accumulation: bit-reproducible totals
    scale = torch.rand(cout, device=DEV) + 0.5
    shift = torch.randn(cout, device=DEV)
    res = torch.randn(n, h, w, cout, device=DEV).to(torch.bfloat16)
    ref = F.conv2d(x, wt, None, 1, 1).permute(0, 2, 3, 1).to(DEV)
    for after in (False, True):
        out = ops.conv_fwd(xd, wpk, cout, 3, 1, 1, scale=scale, shift=shift, residual=res, act=ops.ACT_LEAKY,
                           slope=0.1, res_after_act=after).float()
        z = ref * scale + shift
        want = F.leaky_relu(z, 0.1) + res.float() if after else F.leaky_relu(z + res.float(), 0.1)
        assert float((out - want).abs().max()) <= 2e-2 * float(want.abs().max())


def test_nchw_fp32_head_output():
    n, h, w, cin, cout = 2, 20, 20, 128, 7
    g = torch.Generator().manual_seed(3)
    x = _bf(torch.randn(n, cin, h, w, generator=g))
    wt = _bf(torch.randn(cout, cin, 1, 1, generator=g) * 0.1)
    bias = torch.randn(cout)
    out = ops.conv_fwd(ops.nchw_to_nhwc(x.to(DEV)), ops.pack_weights(wt.to(DEV), False), cout, 1, 1, 0,
                       out_dtype=torch.float32, shift=bias.to(DEV), nchw_out=True)
    assert out.shape == (n, cout, h, w)
    assert torch.allclose(out.cpu(), F.conv2d(x, wt, bias), rtol=1e-4, atol=1e-4)


def test_bn_pool_upsample_against_torch():
    n, h, w, c = 3, 12, 12, 64
    g = torch.Generator().manual_seed(11)
    y = _bf(torch.randn(n, c, h, w, generator=g) * 2 + 0.3)
    gamma, beta = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g)
    da = _bf(torch.randn(n, c, h, w, generator=g))
    yr = y.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rm, rv = torch.zeros(c), torch.ones(c)
    a_ref = F.leaky_relu(F.batch_norm(yr, rm, rv, gr, br, True, 0.1, 1e-5), 0.1)
    a_ref.backward(da)

    yd = ops.nchw_to_nhwc(y.to(DEV))
    stats = ops.stats_encode(torch.stack([yd.float().sum((0, 1, 2)), (yd.float() ** 2).sum((0, 1, 2))]).flatten())
    f = lambda k: torch.empty(k, device=DEV)
    scale, shift, mean, rstd, coef = f(c), f(c), f(c), f(c), f(3 * c)
    rmd, rvd = torch.zeros(c, device=DEV), torch.ones(c, device=DEV)
    ops.bn_finalize(stats, n * h * w, gamma.to(DEV), beta.to(DEV), None, 1e-5, 0.1, rmd, rvd, scale, shift, mean, rstd)
    assert torch.allclose(rmd.cpu(), rm, rtol=1e-4, atol=1e-5) and torch.allclose(rvd.cpu(), rv, rtol=1e-4, atol=1e-5)
    a = ops.bn_apply_act(yd, scale, shift, ops.ACT_LEAKY, 0.1)
    assert float((ops.nhwc_to_nchw(a, c).cpu() - a_ref.detach()).abs().max()) < 3e-2
    dad = ops.nchw_to_nhwc(da.to(DEV))
    sums = ops.bn_bwd_reduce(dad, yd, None, scale, shift, mean, rstd, ops.ACT_LEAKY, 0.1)
    dg, db = f(c), f(c)
    ops.bn_bwd_finalize(sums, gamma.to(DEV), rstd, n * h * w, coef, dg, db)
    dy = ops.bn_bwd_apply(dad, yd, None, scale, shift, mean, rstd, coef, ops.ACT_LEAKY, 0.1)
    assert torch.allclose(dg.cpu(), gr.grad, rtol=2e-2, atol=2e-2)
    assert torch.allclose(db.cpu(), br.grad, rtol=2e-2, atol=2e-2)
    assert float((ops.nhwc_to_nchw(dy, c).cpu() - yr.grad).abs().max()) < 2e-2 * float(yr.grad.abs().max()) + 1e-3

    # max-pool (both variants) and upsample, forward + backward
    xp = _bf(torch.randn(n, c, h, w, generator=g)).requires_grad_(True)
    xd = ops.nchw_to_nhwc(xp.detach().to(DEV))
    for stride in (2, 1):
        ref = F.max_pool2d(F.pad(xp, (0, 1, 0, 1)) if stride == 1 else xp, 2, stride)
        gy = _bf(torch.randn(ref.shape, generator=g))
        xp.grad = None
        ref.backward(gy)
        out = ops.maxpool_fwd(xd, stride)
        assert torch.equal(ops.nhwc_to_nchw(out, c).cpu(), ref.detach())
        dx = ops.maxpool_bwd(xd, ops.nchw_to_nhwc(gy.to(DEV)), stride)
        assert float((ops.nhwc_to_nchw(dx, c).cpu() - xp.grad).abs().max()) < 2e-2
    up = ops.upsample_fwd(xd)
    assert torch.equal(ops.nhwc_to_nchw(up, c).cpu(), F.interpolate(xp.detach(), scale_factor=2, mode="nearest"))
    gu = _bf(torch.randn(n, c, 2 * h, 2 * w, generator=g))
    dxu = ops.upsample_bwd(ops.nchw_to_nhwc(gu.to(DEV)))
    want = gu.view(n, c, h, 2, w, 2).sum((3, 5))
    assert float((ops.nhwc_to_nchw(dxu, c).cpu() - want).abs().max()) < 3e-2


IMAGE_CASES = [  # N, H, W, k, Cout  (3-channel image layers: CVC-YOLOv3 conv_0, the RektNet stem)
    (2, 32, 64, 3, 32),
    (3, 21, 45, 3, 32),     # ragged: partial tiles in both directions
    (2, 80, 80, 7, 16),
    (2, 19, 70, 7, 16),
    (5, 8, 32, 3, 16),
]


@pytest.mark.parametrize("case", IMAGE_CASES)
def test_image_conv_without_patch_matrix(case):
    """b200cv_conv_image_fwd / _wgrad (halo tile in shared memory + mma.sync) against torch fp32 on the same
    bf16-rounded operands: output, fused BatchNorm statistics (bit-reproducible), folded affine + activation, and the
    flat packed weight gradient."""
    n, h, w, k, cout = case
    pad = (k - 1) // 2
    assert ops.use_image_path(3, k, 1, pad, 1, cout)
    g = torch.Generator().manual_seed(sum(case))
    x = torch.rand(n, 3, h, w, generator=g)
    wt = _bf(torch.randn(cout, 3, k, k, generator=g) * 0.2).requires_grad_(True)
    y_ref = F.conv2d(_bf(x), wt, None, 1, pad)
    dy = _bf(torch.randn(y_ref.shape, generator=g))
    y_ref.backward(dy)
    kp = ops.flat_k(3, k)
    flat = torch.zeros(cout, 1, kp)
    flat[:, 0, :3 * k * k] = wt.detach().permute(0, 2, 3, 1).reshape(cout, -1)
    flat = flat.to(DEV).to(torch.bfloat16)
    xd = x.to(DEV)
    stats = ops.stats_buffer(cout, DEV)
    y = ops.conv_image_fwd(xd, flat, cout, k, pad, stats=stats)
    assert y.shape == (n, h, w, ops.pad_channels(cout))
    got = ops.nhwc_to_nchw(y, cout).cpu()
    assert float((got - y_ref.detach()).abs().max()) <= 1e-2 * float(y_ref.abs().max())
    yf = y.float()[..., :cout]
    tot = ops.stats_value(stats).float()
    assert torch.allclose(tot[:cout], yf.sum((0, 1, 2)), rtol=1e-3, atol=1e-2)
    assert torch.allclose(tot[cout:], (yf * yf).sum((0, 1, 2)), rtol=1e-3, atol=1e-2)
    again = ops.stats_buffer(cout, DEV)
    y2 = ops.conv_image_fwd(xd, flat, cout, k, pad, stats=again)
    assert torch.equal(y, y2) and torch.equal(stats.sum(0), again.sum(0))
    scale, shift = torch.rand(cout, device=DEV) + 0.5, torch.randn(cout, device=DEV)
    out = ops.conv_image_fwd(xd, flat, cout, k, pad, scale=scale, shift=shift, act=ops.ACT_LEAKY, slope=0.1)
    want = F.leaky_relu(y_ref.detach().to(DEV) * scale[None, :, None, None] + shift[None, :, None, None], 0.1)
    assert float((ops.nhwc_to_nchw(out, cout) - want).abs().max()) <= 1.5e-2 * float(want.abs().max())
    dwp = torch.zeros(cout, 1, kp, device=DEV)
    ops.conv_image_wgrad(xd, ops.nchw_to_nhwc(dy.to(DEV)), cout, k, pad, 1, dwp)
    want_dw = wt.grad.permute(0, 2, 3, 1).reshape(cout, -1)
    assert float((dwp[:, 0, :3 * k * k].cpu() - want_dw).abs().max()) <= 2e-3 * float(want_dw.abs().max())
    assert float(dwp[:, 0, 3 * k * k:].abs().max()) == 0.0 if kp > 3 * k * k else True
    # weight gradient with the layer's own BatchNorm + activation backward applied on the fly == the two-pass form
    # (bn_bwd_stats_apply writes dy in bf16, conv_image_wgrad reads it): same rounding points, same constants
    c = ops.pad_channels(cout)
    da = torch.randn(n, h, w, c, generator=g).to(DEV).to(torch.bfloat16)
    gamma = (torch.rand(cout, generator=g) + 0.5).to(DEV)
    bscale, bshift = (torch.rand(cout, generator=g) + 0.5).to(DEV), (torch.randn(cout, generator=g) * 0.3).to(DEV)
    mean, rstd = (torch.randn(cout, generator=g) * 0.1).to(DEV), (torch.rand(cout, generator=g) + 0.5).to(DEV)
    parts = ops.bn_bwd_reduce(da, y, None, bscale, bshift, mean, rstd, ops.ACT_LEAKY, 0.1)
    coef_a, coef_b = torch.zeros(3 * cout, device=DEV), torch.zeros(3 * cout, device=DEV)
    dg_a, db_a, dg_b, db_b = (torch.zeros(cout, device=DEV) for _ in range(4))
    dy2 = ops.bn_bwd_stats_apply(parts, n * h * w, gamma, coef_a, dg_a, db_a, da, y, bscale, bshift, mean, rstd,
                                 ops.ACT_LEAKY, 0.1)
    two = ops.conv_image_wgrad(xd, dy2, cout, k, pad, 1, torch.zeros(cout, 1, kp, device=DEV))
    one = ops.conv_image_wgrad_bn(xd, da, y, parts, n * h * w, gamma, coef_b, dg_b, db_b, bscale, bshift, mean, rstd,
                                  ops.ACT_LEAKY, 0.1, cout, k, pad, 1, torch.zeros(cout, 1, kp, device=DEV))
    assert torch.equal(coef_a, coef_b) and torch.equal(dg_a, dg_b) and torch.equal(db_a, db_b)
    assert float((one - two).abs().max()) <= 1e-4 * float(two.abs().max())


@pytest.mark.parametrize("case", [(2, 32, 64, 32, 64), (3, 20, 40, 64, 128), (1, 52, 52, 128, 256), (2, 9, 36, 32, 48), (2, 38, 76, 64, 128)])
def test_stride2_data_gradient_as_one_depth_to_space_launch(case):
    """b200cv_conv_dgrad_d2s (one GEMM over the dy grid + 3-D TMA stores in row groups that never cross the end of a
    dy row) equals the four-launch parity-class data gradient and torch on the same bf16-rounded operands."""
    from b200cv.packing import ConvPackSet, GradArena

    n, oh, ow, cin, cout = case
    g = torch.Generator().manual_seed(sum(case))
    conv = torch.nn.Conv2d(cin, cout, 3, 2, 1, bias=False).to(DEV)
    with torch.no_grad():
        conv.weight.copy_(_bf(torch.randn(cout, cin, 3, 3, generator=g) * 0.1))
    packs = ConvPackSet([(conv, True)], DEV, GradArena([conv.weight], DEV), d2s=[conv])
    packs.pack_all(True)
    x = torch.zeros(n, cin, 2 * oh, 2 * ow, requires_grad=True)
    dy = _bf(torch.randn(n, cout, oh, ow, generator=g))
    F.conv2d(x, conv.weight.detach().cpu(), None, 2, 1).backward(dy)
    dyd = ops.nchw_to_nhwc(dy.to(DEV))
    assert ops.d2s_dgrad_ok(cin, 3, 2, 1, ow)
    dx = ops.conv_dgrad_d2s(dyd, packs.wpk_d2s[id(conv)], cin)
    four = ops.conv_dgrad(dyd, packs.wpk_t[id(conv)], cin, 3, 2, 1, 1, (2 * oh, 2 * ow))
    assert dx.shape == four.shape
    ref = x.grad
    assert float((ops.nhwc_to_nchw(dx, cin).cpu() - ref).abs().max()) <= 1.5e-2 * float(ref.abs().max())
    assert float((dx.float() - four.float()).abs().max()) <= 1e-2 * float(ref.abs().max())
    if cin & (cin - 1) == 0:
        # the fused first pass of the BatchNorm backward (sums of dz and dz*xhat per CHANNEL: four column groups of the
        # depth-to-space GEMM share a channel) equals the stand-alone reduction over the stored gradient
        y = torch.randn(n, 2 * oh, 2 * ow, cin, generator=g).to(DEV).to(torch.bfloat16)
        scale, shift = (torch.rand(cin, generator=g) + 0.5).to(DEV), torch.randn(cin, generator=g).to(DEV) * 0.3
        mean, rstd = torch.randn(cin, generator=g).to(DEV) * 0.1, (torch.rand(cin, generator=g) + 0.5).to(DEV)
        sums = ops.stats_buffer(cin, DEV)
        dx2 = ops.conv_dgrad_d2s(dyd, packs.wpk_d2s[id(conv)], cin,
                                 bn_reduce=(y, scale, shift, mean, rstd, ops.ACT_LEAKY, 0.1, sums))
        assert torch.equal(dx2, dx)
        want = ops.stats_value(ops.bn_bwd_reduce(dx, y, None, scale, shift, mean, rstd, ops.ACT_LEAKY, 0.1)).float()
        got = ops.stats_value(sums).float()
        assert torch.allclose(got, want, rtol=2e-3, atol=2e-3 * float(want.abs().max())), (got - want).abs().max()
