"""Pins the oracle: oracle/*.py must reproduce the vectors the REFERENCE produced (tests/golden/,
made by oracle/gen_golden.py importing /root/reference).  CPU only."""
import pytest
import torch

import helpers
from b200cv import cfg_gen
from oracle import rektnet_oracle as RO
from oracle import yolo_oracle as YO


def test_build_targets_all_cases(golden_yolo):
    cases = golden_yolo["build_targets"]
    assert len(cases) >= 20
    for name, c in cases.items():
        got = YO.build_targets(c["targets"].clone(), c["anchors"], c["C"], c["G"], c["G"], 0.5)
        for j, (a, b) in enumerate(zip(got, c["out"])):
            assert a.dtype == b.dtype and a.shape == b.shape, (name, j)
            if a.dtype == torch.uint8:
                assert torch.equal(a, b), (name, j)  # masks / indices: bit-exact
            else:
                assert torch.allclose(a, b, rtol=1e-6, atol=1e-6), (name, j)


@pytest.mark.parametrize("name", ["c1_g13", "c1_g26", "c3_g13", "c80_g13", "c1_g13_saturated"])
def test_yolo_layer(golden_yolo, name):
    g = golden_yolo["yolo_layer"][name]
    gen = torch.Generator().manual_seed(100 + g["tseed"])
    sample = (torch.randn(g["B"], 3 * (5 + g["C"]), g["G"], g["G"], generator=gen) * g["scale"]).requires_grad_(True)
    tg = YO.synth_targets(g["B"], 8, seed=g["tseed"])
    anchors = [YO.VANILLA_ANCHORS[i] for i in (6, 7, 8)]
    loss, parts = YO.yolo_layer(sample, tg, anchors, g["C"], 416, 0.5, 2.0, 1.6, 0.1, 25.0)
    loss.backward()
    assert torch.allclose(loss.detach(), g["loss"], rtol=1e-6)
    assert torch.allclose(parts, g["parts"], rtol=1e-6)
    assert int((sample.grad != 0).sum()) == g["grad_nnz"]
    assert abs(float(sample.grad.double().abs().sum()) - g["grad_abs_sum"]) <= 1e-6 * g["grad_abs_sum"]
    if "grad" in g:
        assert torch.allclose(sample.grad, g["grad"], rtol=1e-5, atol=1e-8)
    det = YO.yolo_layer(sample.detach(), None, anchors, g["C"], 416, 0.5, 2.0, 1.6, 0.1, 25.0)
    assert torch.allclose(det[:, ::37], g["det_rows"], rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("name", ["tiny_128", "tiny_416", "full_128", "tiny_128_c80"])
def test_darknet(cfg_dir, golden_yolo, name):
    g = golden_yolo["darknet"][name]
    model, path = helpers.make_darknet(cfg_dir, g["cfg"], g["S"], g["C"])
    helpers.assert_digest(model.named_parameters(), g["digest"])
    params = {k: v.detach().clone().requires_grad_(True) for k, v in model.named_parameters()}
    buffers = {k: v.clone() for k, v in model.named_buffers()}
    spec = YO.NetSpec(path)
    x = YO.synth_images(g["B"], g["S"], g["S"], seed=0)
    tg = YO.synth_targets(g["B"], 16, seed=1)
    losses = YO.darknet_forward(spec, params, buffers, x, tg)
    losses[0].sum().backward()
    got = torch.stack([l.detach() for l in losses])
    assert torch.allclose(got, g["losses"], rtol=2e-5), (got, g["losses"])

    class P:  # adapter for helpers.grad_errors
        def __init__(self, t):
            self.grad, self._t = t.grad, t

        def numel(self):
            return self._t.numel()

    errs = helpers.grad_errors([(k, P(v)) for k, v in params.items()], g["grads"])
    assert max(errs.values()) < 1e-3, sorted(errs.items(), key=lambda kv: -kv[1])[:3]
    for k, (s, a) in g["running"].items():
        assert abs(float(buffers[k].double().sum()) - s) <= 1e-5 * (a + 1), k
    with torch.no_grad():
        det = YO.darknet_forward(spec, params, buffers, x, None, training=False)
    assert tuple(det.shape) == g["det_shape"]
    assert torch.allclose(det[:, ::53], g["det_rows"], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("loss_type", ["l2_softargmax", "l2_heatmap", "l1_softargmax"])
@pytest.mark.parametrize("geo", [False, True])
def test_rektnet(golden_rekt, loss_type, geo):
    g = golden_rekt[f"{loss_type}_geo{int(geo)}"]
    import keypoint_net

    torch.manual_seed(17)
    net = keypoint_net.KeypointNet()
    helpers.assert_digest(net.named_parameters(), g["digest"])
    params = {k: v.detach().clone().requires_grad_(True) for k, v in net.named_parameters()}
    buffers = {k: v.clone() for k, v in net.named_buffers()}
    x, thm, tpts = RO.synth_batch(g["B"], seed=0)
    hm, pts = RO.keypointnet_forward(params, buffers, x, training=True)
    loc, geo_l, total = RO.cross_ratio_loss(hm, pts, thm, tpts, loss_type, geo, 0.055, 0.038)
    total.backward()
    assert torch.allclose(loc.detach(), g["loc"], rtol=1e-5)
    assert torch.allclose(geo_l.detach().float(), g["geo"], rtol=1e-5, atol=1e-7)
    assert torch.allclose(total.detach(), g["total"], rtol=1e-5)
    assert torch.allclose(pts.detach(), g["pts"], rtol=1e-5, atol=1e-6)
    assert torch.allclose(hm.detach()[:, :, ::16, ::16], g["hm_rows"], rtol=1e-4, atol=1e-9)

    class P:
        def __init__(self, t):
            self.grad, self._t = t.grad, t

        def numel(self):
            return self._t.numel()

    errs = helpers.grad_errors([(k, P(v)) for k, v in params.items()], g["grads"])
    assert max(errs.values()) < 2e-3, sorted(errs.items(), key=lambda kv: -kv[1])[:3]


def test_rektnet_eval(golden_rekt):
    import keypoint_net

    torch.manual_seed(17)
    net = keypoint_net.KeypointNet()
    params = {k: v.detach() for k, v in net.named_parameters()}
    buffers = {k: v.clone() for k, v in net.named_buffers()}
    x, _, _ = RO.synth_batch(4, seed=0)
    hm, pts = RO.keypointnet_forward(params, buffers, x, training=False)
    assert torch.allclose(pts, golden_rekt["eval"]["pts"], rtol=1e-5, atol=1e-6)
