"""KeypointNet + CrossRatioLoss on the B200 path vs goldens made by the reference."""
import pytest
import torch

import helpers
from oracle import rektnet_oracle as RO

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _net():
    import keypoint_net

    torch.manual_seed(17)
    return keypoint_net.KeypointNet()


def _cos(a, b):
    a, b = a.flatten().double(), b.flatten().double()
    return float((a * b).sum() / (a.norm() * b.norm() + 1e-30))


@pytest.mark.parametrize("loss_type", ["l2_softargmax", "l2_heatmap", "l1_softargmax"])
@pytest.mark.parametrize("geo", [False, True])
def test_rektnet_train_step_vs_reference(golden_rekt, loss_type, geo):
    """Losses/outputs vs the REFERENCE golden (bf16 tolerance); gradients vs the oracle with bf16 storage
    emulated (kernel correctness) and gradient norms vs the fp32 reference."""
    import cross_ratio_loss

    g = golden_rekt[f"{loss_type}_geo{int(geo)}"]
    net = _net()
    helpers.assert_digest(net.named_parameters(), g["digest"])
    cpu_params = {k: v.detach().clone().requires_grad_(True) for k, v in net.named_parameters()}
    cpu_buffers = {k: v.clone() for k, v in net.named_buffers()}
    net = net.to(DEV).train()
    xc, thmc, tptsc = RO.synth_batch(g["B"], seed=0)
    x, thm, tpts = xc.to(DEV), thmc.to(DEV), tptsc.to(DEV)
    hm, pts = net(x)
    loss_fn = cross_ratio_loss.CrossRatioLoss(loss_type, geo, 0.055, 0.038)
    loc, geo_l, total = loss_fn(hm, pts, thm, tpts)
    total.backward()
    assert abs(float(loc) - float(g["loc"])) <= 1e-2 * abs(float(g["loc"]))
    assert abs(float(geo_l) - float(g["geo"])) <= 1e-2 * abs(float(g["geo"])) + 1e-6
    assert abs(float(total) - float(g["total"])) <= 1e-2 * abs(float(g["total"]))
    # random-init heat-maps are almost flat: the soft-argmax is a sensitive statistic of the logits
    assert float((pts.detach().cpu() - g["pts"]).abs().max()) < 5e-2
    ehm, epts = RO.keypointnet_forward(cpu_params, cpu_buffers, xc, True, emulate_bf16=True)
    RO.cross_ratio_loss(ehm, epts, thmc, tptsc, loss_type, geo, 0.055, 0.038)[2].backward()
    assert float((pts.detach().cpu() - epts.detach()).abs().max()) < 4e-2  # run-to-run spread seen: up to 2.6e-2
    assert float((hm.detach().cpu() - ehm.detach()).norm() / ehm.detach().norm()) < 5e-2
    skip = ("conv.bias", "conv1.bias", "conv2.bias", "shortcut_conv.bias", "out.bias")  # ~0 by construction
    for k, p in net.named_parameters():
        if k.endswith(skip):
            continue
        assert _cos(p.grad.cpu(), cpu_params[k].grad) > 0.97, k
        ref = g["grads"][k]["norm"]
        assert abs(float(p.grad.double().norm()) - ref) <= 0.15 * ref + 1e-6, k


def test_unfused_loss_backward_equals_fused(golden_rekt):
    """CrossRatioLoss on tensors that lost the KeypointNet tag takes the un-fused kernels; same gradients
    (up to the run-to-run noise of fp32 atomics in the BN statistics)."""
    import cross_ratio_loss

    x, thm, tpts = (t.to(DEV) for t in RO.synth_batch(4, seed=0))
    grads = []
    for fused in (True, False):
        net = _net().to(DEV).train()
        hm, pts = net(x)
        if not fused:
            hm, pts = hm * 1.0, pts * 1.0
        loss_fn = cross_ratio_loss.CrossRatioLoss("l2_heatmap", True, 0.055, 0.038)
        loss_fn(hm, pts, thm, tpts)[2].backward()
        grads.append({k: p.grad.clone() for k, p in net.named_parameters() if "weight" in k})
    for k in grads[0]:
        assert _cos(grads[0][k], grads[1][k]) > 0.99, k
        assert abs(float(grads[0][k].norm() / grads[1][k].norm()) - 1) < 0.08, k


def test_rektnet_eval_and_onnx_mode(golden_rekt):
    import keypoint_net

    net = _net().to(DEV).eval()
    x, _, _ = RO.synth_batch(4, seed=0)
    with torch.no_grad():
        hm, pts = net(x.to(DEV))
    assert float((pts.cpu() - golden_rekt["eval"]["pts"]).abs().max()) < 5e-3
    assert torch.allclose(hm.sum((2, 3)).cpu(), torch.ones(4, 7), atol=1e-4)
    torch.manual_seed(17)
    raw = keypoint_net.KeypointNet(onnx_mode=True).to(DEV).eval()
    with torch.no_grad():
        logits = raw(x.to(DEV))
    assert logits.shape == (4, 7, 80, 80)
    assert torch.allclose(torch.softmax(logits.view(4, 7, -1), -1).view_as(hm), hm, rtol=1e-3, atol=1e-7)


def test_soft_argmax_delta():
    """A one-hot heat-map at (y=10, x=30) decodes to (0.375, 0.125) -- SURVEY 8a-12."""
    from b200cv.lib import lib, ptr, stream_ptr

    logits = torch.full((1, 1, 80, 80), -1e4, device=DEV)
    logits[0, 0, 10, 30] = 50.0
    vx = torch.linspace(0, 79.0 / 80, 80, device=DEV)
    hm, pts = torch.empty_like(logits), torch.empty(1, 1, 2, device=DEV)
    lib().call("b200cv_kpt_softmax_argmax", ptr(logits), ptr(vx), ptr(vx), ptr(hm), ptr(pts), 1, 80, 80, stream_ptr())
    assert torch.allclose(pts.cpu().view(-1), torch.tensor([0.375, 0.125]), atol=1e-6)


@pytest.mark.parametrize("cin,cout", [(16, 32), (64, 128)])
def test_resnet_block_on_its_own(cin, cout):
    """resnet.ResNet.forward as a callable block (RektNet/resnet.py:22-27): output, parameter gradients and the
    gradient of its INPUT vs the oracle's res_block (pinned to the reference) on the same weights."""
    import resnet

    torch.manual_seed(3)
    blk = resnet.ResNet(cin, cout)
    params = {f"b.{k}": v.detach().clone().requires_grad_(True) for k, v in blk.named_parameters()}
    buffers = {f"b.{k}": v.clone() for k, v in blk.named_buffers()}
    g = torch.Generator().manual_seed(5)
    xc = torch.randn(4, cin, 40, 40, generator=g).requires_grad_(True)
    want = RO.res_block(xc, params, buffers, "b", True)
    up = torch.randn(want.shape, generator=g)
    want.backward(up)
    blk = blk.to(DEV).train()
    x = xc.detach().to(DEV).requires_grad_(True)
    out = blk(x)
    assert out.shape == want.shape and out.dtype == torch.float32
    out.backward(up.to(DEV))
    assert float((out.detach().cpu() - want.detach()).norm() / want.detach().norm()) < 1e-2
    assert _cos(x.grad.cpu(), xc.grad) > 0.995
    for k, p in blk.named_parameters():
        if k.endswith("bias") and "bn" not in k:
            continue  # conv biases: analytically zero gradient under train-mode BN
        assert _cos(p.grad.cpu(), params[f"b.{k}"].grad) > 0.99, k
    for k, b in blk.named_buffers():  # running statistics follow the reference (conv bias included in the mean)
        if b.dtype.is_floating_point:
            assert torch.allclose(b.cpu(), buffers[f"b.{k}"], rtol=2e-2, atol=2e-3), k
    blk.eval()
    y_eval = blk(x.detach())  # eval mode WITHOUT no_grad (RektNet/detect.py:38-39): computed, not differentiable
    assert not y_eval.requires_grad
    want_eval = RO.res_block(xc.detach(), {k: v.detach() for k, v in params.items()},
                             {k: v.cpu() for k, v in (("b." + n, t) for n, t in blk.named_buffers())}, "b", False)
    assert float((y_eval.cpu() - want_eval).norm() / want_eval.norm()) < 1e-2


def test_eval_forward_without_no_grad():
    """`model.eval(); out = model(x)` with grad mode ON must compute: reference RektNet/detect.py:38-39."""
    net = _net().to(DEV).eval()
    x, _, _ = RO.synth_batch(2, seed=0)
    hm, pts = net(x.to(DEV))
    assert not hm.requires_grad and not pts.requires_grad
    with torch.no_grad():
        hm2, pts2 = net(x.to(DEV))
    assert torch.equal(pts, pts2)


def test_cuda_graph_step_matches_eager(monkeypatch):
    """From the third step with a shape the KeypointNet step is replayed from CUDA graphs (forward; backward per loss
    configuration).  With lr = 0 the replayed steps reproduce the eager ones on new inputs: same losses (the forward is
    bit-reproducible), same gradients, same running statistics."""
    import cross_ratio_loss

    res = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("B200CV_CUDA_GRAPH", mode)
        net = _net().to(DEV).train()
        loss_fn = cross_ratio_loss.CrossRatioLoss("l2_heatmap", True, 0.055, 0.038)
        opt = torch.optim.SGD(net.parameters(), lr=0.0)
        hist = []
        for it in range(5):
            x, thm, tpts = (t.to(DEV) for t in RO.synth_batch(4, seed=it))
            opt.zero_grad()
            hm, pts = net(x)
            loss = loss_fn(hm, pts, thm, tpts)
            loss[2].backward()
            opt.step()
            hist.append((torch.stack([l.detach().float().to(DEV) for l in loss]).cpu(), pts.detach().cpu().clone(),
                         {k: p.grad.detach().cpu().clone() for k, p in net.named_parameters()}))
        res[mode] = (hist, {k: v.detach().cpu().clone() for k, v in net.state_dict().items()})
        if mode == "1":
            from b200cv.rektnet_engine import _GraphedRektStep

            steps = [v for v in net.engine()._graphs.values() if isinstance(v, _GraphedRektStep)]
            assert steps and len(steps[0].bwd) == 1  # one backward graph for the one loss configuration
    for (la, pa, ga), (lb, pb, gb) in zip(res["0"][0], res["1"][0]):
        assert torch.equal(la, lb) and torch.equal(pa, pb)
        for k in ga:
            if k == "out.bias":
                continue  # analytically zero (softmax Jacobian): what is left is the rounding noise of a float reduction
            assert torch.allclose(ga[k], gb[k], rtol=1e-4, atol=1e-6 * float(ga[k].abs().max()) + 1e-12), k
    for k, v in res["0"][1].items():
        assert torch.equal(v, res["1"][1][k]), k
