"""YOLO head kernels vs the reference goldens and the oracle (indices/masks bit-exact)."""
import pytest
import torch

from oracle import yolo_oracle as YO

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_build_targets_bit_exact_all_golden_cases(golden_yolo):
    from utils.utils import build_targets

    for name, c in golden_yolo["build_targets"].items():
        got = build_targets(c["targets"].to(DEV), c["anchors"].to(DEV), 3, c["C"], c["G"], c["G"], 0.5)
        for j, (a, b) in enumerate(zip(got, c["out"])):
            assert a.dtype == b.dtype and tuple(a.shape) == tuple(b.shape), (name, j)
            if a.dtype == torch.uint8:
                assert torch.equal(a.cpu(), b), (name, j)
            else:
                assert torch.allclose(a.cpu(), b, rtol=1e-5, atol=1e-6), (name, j)


@pytest.mark.parametrize("B,T,G,C", [(8, 16, 13, 1), (8, 16, 52, 1), (4, 16, 19, 80), (16, 4, 76, 1), (3, 1, 26, 2)])
def test_build_targets_random_vs_oracle(B, T, G, C):
    from utils.utils import build_targets

    tg = YO.synth_targets(B, T, seed=B * 100 + G)
    tg[0] = 0  # an image without any target -> phantom positive
    if T > 2:
        tg[1, 1] = tg[1, 0]  # exact duplicate -> collision
    tg[:, :, 0] = torch.randint(0, C, (B, T)).float() * (tg[:, :, 1:].sum(-1) > 0)
    anchors = YO.scaled_anchors([YO.VANILLA_ANCHORS[i] for i in (3, 4, 5)], 416 / G)
    want = YO.build_targets(tg.clone(), anchors, C, G, G, 0.5)
    got = build_targets(tg.to(DEV), anchors.to(DEV), 3, C, G, G, 0.5)
    for j, (a, b) in enumerate(zip(got, want)):
        if a.dtype == torch.uint8:
            assert torch.equal(a.cpu(), b), j
        else:
            assert torch.allclose(a.cpu(), b, rtol=1e-5, atol=1e-6), j


@pytest.mark.parametrize("name", ["c1_g13", "c1_g26", "c3_g13", "c80_g13", "c1_g13_saturated"])
def test_yolo_layer_vs_golden(golden_yolo, name):
    import models

    g = golden_yolo["yolo_layer"][name]
    gen = torch.Generator().manual_seed(100 + g["tseed"])
    sample = (torch.randn(g["B"], 3 * (5 + g["C"]), g["G"], g["G"], generator=gen) * g["scale"])
    sample = sample.to(DEV).requires_grad_(True)
    tg = YO.synth_targets(g["B"], 8, seed=g["tseed"]).to(DEV)
    anchors = [YO.VANILLA_ANCHORS[i] for i in (6, 7, 8)]
    layer = models.YOLOLayer(anchors, g["C"], 416, 416, 0.5, "leaky", 2.0, 1.6, 0.1, 25.0)
    loss, parts = layer(sample, tg)
    loss.backward()
    assert torch.allclose(loss.detach().cpu(), g["loss"], rtol=1e-5)
    assert torch.allclose(parts.cpu(), g["parts"], rtol=1e-5)
    grad = sample.grad.cpu()
    assert int((grad != 0).sum()) == g["grad_nnz"]
    assert abs(float(grad.double().abs().sum()) - g["grad_abs_sum"]) <= 1e-5 * g["grad_abs_sum"]
    if "grad" in g:
        assert torch.allclose(grad, g["grad"], rtol=1e-4, atol=1e-8)
    det = layer(sample.detach())
    assert torch.allclose(det.cpu()[:, ::37], g["det_rows"], rtol=1e-5, atol=1e-5)


def test_yolo_loss_nhwc_bf16_engine_layout_matches_strided():
    """The engine's channel-contiguous kernel (bf16 dlogits incl. zero class/pad channels) agrees
    with the strided fp32 kernel used by the public YOLOLayer."""
    from b200cv import yolo_ops

    B, C, G = 4, 80, 13
    nch = 3 * (5 + C)
    gen = torch.Generator().manual_seed(5)
    z_nchw = torch.randn(B, nch, G, G, generator=gen).to(DEV)
    z_nhwc = torch.zeros(B, G, G, 256, device=DEV)
    z_nhwc[..., :nch] = z_nchw.permute(0, 2, 3, 1)
    tg = YO.synth_targets(B, 8, seed=2).to(DEV)
    sa = yolo_ops.scaled_anchors([YO.VANILLA_ANCHORS[i] for i in (6, 7, 8)], 32.0, DEV)
    yt = yolo_ops.yolo_targets(tg, sa, G, G, 0.5)
    consts = (2.0, 1.6, 0.1, 25.0)
    s1 = torch.zeros(6, dtype=torch.float64, device=DEV)
    s2 = torch.zeros(6, dtype=torch.float64, device=DEV)
    d1 = torch.zeros_like(z_nchw)
    d2 = torch.full((B, G, G, 256), 7.0, device=DEV).to(torch.bfloat16)
    gs = torch.tensor([0.5], device=DEV)
    yolo_ops.yolo_loss(z_nchw, True, yt, C, consts, sums=s1, dlogits=d1, gscale=gs)
    yolo_ops.yolo_loss(z_nhwc, False, yt, C, consts, sums=s2, dlogits=d2, gscale=gs)
    assert torch.allclose(s1, s2, rtol=1e-5)
    # the engine's two-kernel form (cell kernel + streaming expansion) gives the same dense gradient
    s3 = torch.zeros(6, dtype=torch.float64, device=DEV)
    yolo_ops.yolo_loss_cells(z_nhwc, False, yt, C, consts, sums=s3)
    d3 = yolo_ops.yolo_head_grad(z_nhwc, yt, C, consts, gs)
    assert torch.allclose(s1, s3, rtol=1e-5) and torch.equal(d3, d2)
    assert float(d2[..., nch:].float().abs().max()) == 0.0
    want = d1.permute(0, 2, 3, 1)
    assert torch.allclose(d2[..., :nch].float(), want, rtol=1e-2, atol=1e-7)
    assert int((d2 != 0).sum()) == int((want.to(torch.bfloat16) != 0).sum())


def test_yolo_loss_full_size_properties():
    """BASELINE size (416^2, bs64, C=80, three scales): size-independent properties."""
    from b200cv import yolo_ops

    B, C, T = 64, 80, 16
    tg = YO.synth_targets(B, T, seed=1).to(DEV)
    consts = (2.0, 1.6, 0.1, 25.0)
    for G, mask in ((13, (6, 7, 8)), (26, (3, 4, 5)), (52, (0, 1, 2))):
        z = torch.randn(B, G, G, 256, device=DEV)
        sa = yolo_ops.scaled_anchors([YO.VANILLA_ANCHORS[i] for i in mask], 416 / G, DEV)
        yt = yolo_ops.yolo_targets(tg, sa, G, G, 0.5)
        nm, nf = yt.counts.tolist()
        owner = yt.owner
        assert nm == int((owner >= 0).sum()) and nm >= B
        ign_cells = yt.ign.bool()
        assert nf == int(((owner < 0) & ~ign_cells[None, None]).sum())
        out7 = torch.zeros(7, device=DEV)
        sums = torch.zeros(6, dtype=torch.float64, device=DEV)
        d_a = torch.empty(B, G, G, 256, device=DEV, dtype=torch.bfloat16)
        d_b = torch.empty_like(d_a)
        yolo_ops.yolo_loss(z, False, yt, C, consts, sums=sums, dlogits=d_a, gscale=torch.tensor([1.0], device=DEV))
        yolo_ops.yolo_loss(z, False, yt, C, consts, dlogits=d_b, gscale=torch.tensor([2.0], device=DEV))
        yolo_ops.yolo_loss_finalize(sums, yt, consts, out7)
        assert torch.isfinite(out7).all()
        assert abs(float(out7[0]) - float(out7[1:].sum())) <= 1e-4 * float(out7[0])  # total = sum of parts
        assert torch.equal((d_a.float() * 2).to(torch.bfloat16), d_b)  # linear in the upstream gradient
        nz = d_a.view(B, G, G, 256)[..., :255].view(B, G, G, 3, 85)
        assert float(nz[..., 5:].float().abs().max()) == 0.0  # class channels get exactly zero
        xywh_nz = (nz[..., :4] != 0).any(-1)
        assert int(xywh_nz.sum()) <= nm and int(xywh_nz.sum()) >= nm - 2  # only object cells (exact-zero diffs aside)


def test_decode_three_scales_row_order():
    import models

    B, C = 2, 1
    for G, mask in ((13, (6, 7, 8)), (26, (3, 4, 5))):
        anchors = [YO.VANILLA_ANCHORS[i] for i in mask]
        sample = torch.randn(B, 3 * (5 + C), G, G, generator=torch.Generator().manual_seed(G))
        layer = models.YOLOLayer(anchors, C, 416, 416, 0.5, "leaky", 2.0, 1.6, 0.1, 25.0)
        got = layer(sample.to(DEV)).cpu()
        want = YO.yolo_layer(sample, None, anchors, C, 416, 0.5, 2.0, 1.6, 0.1, 25.0)
        assert torch.allclose(got, want, rtol=1e-5, atol=1e-5)
