"""Fused optimizer steps against torch.optim on the same device tensors."""
import pytest
import torch

from b200cv import optim as boptim

pytestmark = pytest.mark.gpu

SHAPES = [(64, 32, 3, 3), (255,), (1024, 512, 1, 1), (7,), (33, 5), (3,), (100003,)]


def _params(seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.nn.Parameter(torch.randn(*s, generator=g).cuda()) for s in SHAPES]


def _set_grads(ps, step, arena=None):
    g = torch.Generator().manual_seed(100 + step)
    off = 0
    for p in ps:
        gr = torch.randn(p.shape, generator=g).cuda() * 0.1
        if arena is None:
            p.grad = gr
        else:  # gradients as (possibly 4-byte-aligned only) views of one flat arena, like the engine's GradArena
            v = arena[off:off + p.numel()].view_as(p)
            v.copy_(gr)
            p.grad = v
            off += p.numel()


@pytest.mark.parametrize("wd", [0.0, 5e-4])
@pytest.mark.parametrize("use_arena", [False, True])
def test_fused_adam_matches_torch(wd, use_arena):
    a, b = _params(0), _params(0)
    ref = torch.optim.Adam(a, lr=1e-2, weight_decay=wd)
    mine = boptim.FusedAdam(b, lr=1e-2, weight_decay=wd)
    sched_r = torch.optim.lr_scheduler.StepLR(ref, step_size=1, gamma=0.95)   # CVC-YOLOv3/train.py:199
    sched_m = torch.optim.lr_scheduler.StepLR(mine, step_size=1, gamma=0.95)
    arena = torch.zeros(sum(p.numel() for p in b) + 1, device="cuda")[1:] if use_arena else None
    for step in range(6):
        _set_grads(a, step)
        _set_grads(b, step, arena)
        ref.step()
        mine.step()
        sched_r.step()
        sched_m.step()
    for p, q in zip(a, b):
        assert torch.allclose(p, q, rtol=2e-5, atol=1e-6), p.shape
        assert torch.allclose(ref.state[p]["exp_avg"], mine.state[q]["exp_avg"], rtol=1e-5, atol=1e-7)
        assert torch.allclose(ref.state[p]["exp_avg_sq"], mine.state[q]["exp_avg_sq"], rtol=1e-5, atol=1e-9)
        assert int(mine.state[q]["step"]) == 6
    sd = mine.state_dict()
    assert set(sd["state"][0]) == {"step", "exp_avg", "exp_avg_sq"}


@pytest.mark.parametrize("momentum,wd", [(0.0, 0.0), (0.9, 0.0), (0.9, 5e-4)])
def test_fused_sgd_matches_torch(momentum, wd):
    a, b = _params(1), _params(1)
    ref = torch.optim.SGD(a, lr=0.05, momentum=momentum, weight_decay=wd)
    mine = boptim.FusedSGD(b, lr=0.05, momentum=momentum, weight_decay=wd)
    for step in range(5):
        _set_grads(a, step)
        _set_grads(b, step)
        ref.step()
        mine.step()
    for p, q in zip(a, b):
        assert torch.allclose(p, q, rtol=1e-5, atol=1e-6), p.shape
        if momentum:
            assert torch.allclose(ref.state[p]["momentum_buffer"], mine.state[q]["momentum_buffer"], rtol=1e-5,
                                  atol=1e-7)


def test_fused_adam_drives_the_rektnet_step():
    """RektNet/train_eval.py:59-72 with the fused step on the engine's own gradients (views of its flat arena): the
    update of every iteration equals torch.optim.Adam's applied to the same gradients and state.  (Whole trajectories
    of two separate runs are not comparable: the step is reproducible only to ~1e-3, see tools/determinism_probe.py,
    and Adam's normalised update amplifies that; descent is asserted on the Darknet step in test_gpu_darknet.py.)"""
    import cross_ratio_loss
    import keypoint_net
    from oracle import rektnet_oracle as RO

    torch.manual_seed(5)
    net = keypoint_net.KeypointNet().cuda().train()
    params = list(net.parameters())
    opt = boptim.FusedAdam(params, lr=1e-3, weight_decay=1e-4)
    loss_fn = cross_ratio_loss.CrossRatioLoss("l2_heatmap", True, 0.055, 0.038)
    x, thm, tpts = (t.cuda() for t in RO.synth_batch(8, seed=0))
    shadow = [torch.nn.Parameter(p.detach().clone()) for p in params]  # updated by torch's Adam, same gradients
    ref = torch.optim.Adam(shadow, lr=1e-3, weight_decay=1e-4)
    for it in range(4):
        opt.zero_grad()
        hm, pts = net(x)
        _, _, loss = loss_fn(hm, pts, thm, tpts)
        loss.backward()
        for s, p in zip(shadow, params):
            s.grad = p.grad.detach().clone()
        ref.step()
        opt.step()
        for s, p in zip(shadow, params):
            assert torch.allclose(s, p, rtol=2e-5, atol=1e-6), it
        assert torch.isfinite(loss.detach())
