"""Executor of RektNet's KeypointNet (+ CrossRatioLoss) on the B200 kernels.

``keypoint_net.KeypointNet`` keeps its fp32 nn.Conv2d / nn.BatchNorm2d parameters (same state_dict
names as the reference) but forward/backward run here: tcgen05 convolutions over NHWC bf16
(7x7 stem, dilated 3x3, 3x3, 1x1), two-branch BN+ReLU apply for the residual blocks, a head conv that
writes fp32 NCHW logits, one softmax+soft-argmax kernel, and ONE fused kernel from the loss to the
head-logit gradient.

Fusion across the module boundary: KeypointNet.forward tags the (hm, pts) it returns with a handle.
CrossRatioLoss.forward that sees the tag defers its backward to that handle, so the net's backward
launches a single kernel: loss gradient -> soft-argmax -> softmax Jacobian -> dlogits.  Untagged
tensors take the un-fused kernels (b200cv_kpt_loss_bwd).
"""
from __future__ import annotations

import os

from typing import Optional

import torch

from . import ops
from .lib import lib, ptr, require_cuda, stream_ptr
from .packing import ConvPackSet, GradArena
from .parallel import allreduce_gradients

BN_EPS = 1e-5
BN_MOMENTUM = 0.1
LOSS_TYPES = {"l2_softargmax": 0, "l2_sm": 0, "l2_heatmap": 1, "l2_hm": 1, "l1_softargmax": 2, "l1_sm": 2}


class _ConvBN:
    """A conv followed by a train-mode BN; owns the per-channel scratch vectors."""

    def __init__(self, conv, bn):
        self.conv, self.bn = conv, bn
        self.k, self.pad, self.dil = conv.kernel_size[0], conv.padding[0], conv.dilation[0]
        self.cin, self.cout = conv.in_channels, conv.out_channels
        self.vec_dev = None

    def vecs(self, dev):
        if self.vec_dev != dev:
            c = self.cout
            f = lambda n: torch.empty(n, dtype=torch.float32, device=dev)
            self.stats = ops.stats_buffer(c, dev)
            self.scale, self.shift, self.mean, self.rstd, self.coef = f(c), f(c), f(c), f(c), f(3 * c)
            self.vec_dev = dev

    def fwd_train(self, x):
        self.vecs(x.device)
        self._eval_key = None  # bn_finalize updates the running statistics behind torch's version counters
        self.stats.zero_()
        if x.dtype == torch.float32:  # the stem on the NCHW image itself (csrc/conv_image.cu)
            y = ops.conv_image_fwd(x, self.wpk, self.cout, self.conv.kernel_size[0], self.conv.padding[0],
                                   stats=self.stats)
        else:
            y = ops.conv_fwd(x, self.wpk, self.cout, self.k, 1, self.pad, self.dil, stats=self.stats)
        count = y.numel() // y.shape[-1]
        # the conv bias cancels inside a train-mode BN: it is left out of y and only enters running_mean
        ops.bn_finalize(self.stats, count, self.bn.weight, self.bn.bias, self.conv.bias, BN_EPS, BN_MOMENTUM,
                        self.bn.running_mean, self.bn.running_var, self.scale, self.shift, self.mean, self.rstd)
        if self.bn.num_batches_tracked is not None:
            self.bn.num_batches_tracked += 1
        return y

    def eval_affine(self):
        """Inference BN (+ conv bias) folded into the conv epilogue; cached until a source tensor changes."""
        bn, cb = self.bn, self.conv.bias
        key = (bn.weight._version, bn.bias._version, bn.running_mean._version, bn.running_var._version, cb._version,
               bn.weight.data_ptr(), bn.running_mean.data_ptr())
        if getattr(self, "_eval_key", None) != key:
            scale = bn.weight.detach() * torch.rsqrt(bn.running_var + BN_EPS)
            self._eval_affine = (scale, bn.bias.detach() + (cb.detach() - bn.running_mean) * scale)
            self._eval_key = key
        return self._eval_affine

    def reduce_spec(self, y, act):
        """(bn_reduce tuple for ops.conv_dgrad, zeroed partial sums): lets the data gradient that PRODUCES this
        layer's dL/da also accumulate its BN-backward sums (no separate pass over da and y)."""
        parts = ops.stats_buffer(self.cout, y.device)
        return (y, self.scale, self.shift, self.mean, self.rstd, act, 0.0, parts), parts

    def bwd(self, x_in, y, da, aout, act, gview, packs, dx_out=None, want_dx=True, parts=None, bn_reduce=None,
            side=None):
        """BN+act backward then wgrad (+ dgrad).  Returns dx (or None).  `parts`: BN-backward sums already produced by
        the dgrad that wrote `da`; `bn_reduce`: fused reduction for the layer whose dL/da this dgrad writes."""
        count = y.numel() // y.shape[-1]
        if parts is None:
            parts = ops.bn_bwd_reduce(da, y, aout, self.scale, self.shift, self.mean, self.rstd, act, 0.0)
        if x_in.dtype == torch.float32 and not want_dx and aout is None:
            # the stem on the image: its weight gradient applies the BatchNorm + ReLU backward itself (no dy tensor)
            gview[id(self.conv.bias)].zero_()
            ops.conv_image_wgrad_bn(x_in, da, y, parts, count, self.bn.weight, self.coef, gview[id(self.bn.weight)],
                                    gview[id(self.bn.bias)], self.scale, self.shift, self.mean, self.rstd, act, 0.0,
                                    self.cout, self.conv.kernel_size[0], self.conv.padding[0], 1,
                                    packs.dwp[id(self.conv)])
            return None
        ops.bn_bwd_finalize(parts, self.bn.weight, self.rstd, count, self.coef, gview[id(self.bn.weight)],
                            gview[id(self.bn.bias)])
        dy = ops.bn_bwd_apply(da, y, aout, self.scale, self.shift, self.mean, self.rstd, self.coef, act, 0.0)
        return self.finish(x_in, dy, gview, packs, dx_out, want_dx, bn_reduce, side)

    def finalize_bwd(self, parts, count, gview):
        """BN-backward sums -> d(gamma), d(beta) and the per-channel constants of the apply pass."""
        ops.bn_bwd_finalize(parts, self.bn.weight, self.rstd, count, self.coef, gview[id(self.bn.weight)],
                            gview[id(self.bn.bias)])

    def finish(self, x_in, dy, gview, packs, dx_out=None, want_dx=True, bn_reduce=None, side=None):
        """wgrad (+ dgrad) from the conv-output gradient dy.  With a `_SideQueue` the weight gradient is queued on the
        side stream BEHIND the data gradient: it is off the dependency chain, so it fills the SMs the persistent dgrad
        leaves idle in its last wave and runs under the HBM-bound BatchNorm passes that follow."""
        gview[id(self.conv.bias)].zero_()  # analytically zero under train-mode BN

        def wgrad():
            if x_in.dtype == torch.float32:  # the stem: x_in is the NCHW image
                ops.conv_image_wgrad(x_in, dy, self.cout, self.conv.kernel_size[0], self.conv.padding[0], 1,
                                     packs.dwp[id(self.conv)])
            else:
                ops.conv_wgrad(x_in, dy, self.cout, self.k, 1, self.pad, self.dil, out=packs.dwp[id(self.conv)])

        if side is None:
            wgrad()
        else:
            ready = side.mark()
        dx = None
        if want_dx:
            dx = ops.conv_dgrad(dy, self.wpk_t, self.cin, self.k, 1, self.pad, self.dil,
                                (x_in.shape[1], x_in.shape[2]), out=dx_out, residual=dx_out, bn_reduce=bn_reduce)
        if side is not None:
            side.run(ready, wgrad, x_in, dy)
        return dx


class _SideQueue:
    """A second stream for work that is off the critical path of the backward pass (weight gradients).  Operands are
    kept alive until `join()` so the caching allocator cannot hand their memory out while the side stream reads it."""

    def __init__(self, dev):
        self.main = torch.cuda.current_stream(dev)
        self.side = torch.cuda.Stream(device=dev)
        self.keep = []

    def begin(self):
        self.main = torch.cuda.current_stream(self.side.device)
        self.side.wait_stream(self.main)

    def mark(self):
        ev = torch.cuda.Event()
        ev.record(self.main)
        return ev

    def run(self, ready, fn, *operands):
        self.side.wait_event(ready)
        with torch.cuda.stream(self.side):
            fn()
        self.keep.append(operands)

    def join(self):
        self.main.wait_stream(self.side)
        self.keep.clear()


class _HeadHandle:
    """Links the (hm, pts) returned by KeypointNet.forward with a CrossRatioLoss applied to them."""

    def __init__(self):
        self.pending = None  # dict set by the fused loss backward


class RektNetEngine:
    def __init__(self, model):
        self.model = model
        self.stem = _ConvBN(model.conv, model.bn)
        self.stem.k, self.stem.pad, self.stem.dil = 1, 0, 1  # explicit-im2col stem: a 1x1 conv over the patches
        self.blocks = []
        for r in (model.res1, model.res2, model.res3, model.res4):
            self.blocks.append((_ConvBN(r.conv1, r.bn1), _ConvBN(r.conv2, r.bn2), _ConvBN(r.shortcut_conv, r.shortcut_bn)))
        self.params = list(model.parameters())
        self.split = ops.default_split()  # "fp32" parity mode (split bf16x3 operands), see ops.py
        self._lin = None
        self._arena = None
        self._packs = None

    @property
    def precision(self) -> str:
        return "fp32" if self.split else "bf16"

    def set_precision(self, name: str):
        if name not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        if (name == "fp32") != self.split:
            self.split = name == "fp32"
            self._arena = self._packs = None
        return self

    def _conv_list(self):
        convs = [(self.stem.conv, False)]
        for b in self.blocks:
            convs += [(c.conv, True) for c in b]
        convs.append((self.model.out, True))
        return convs

    def _setup(self, dev):
        if self._arena is None or self._arena.flat.device != dev:
            self._arena = GradArena(self.params, dev)
            self._packs = ConvPackSet(self._conv_list(), dev, self._arena, flat=[self.stem.conv], split=self.split)
            for c in self._all():
                c.wpk, c.wpk_t = self._packs.wpk[id(c.conv)], self._packs.wpk_t[id(c.conv)]

    def _coords(self, dev, h, w):
        if self._lin is None or self._lin[0].device != dev or self._lin[0].numel() != w or self._lin[1].numel() != h:
            vy = torch.linspace(0, (h - 1.0) / h, h, dtype=torch.float32, device=dev)  # keypoint_net.py:52-53
            vx = torch.linspace(0, (w - 1.0) / w, w, dtype=torch.float32, device=dev)
            self._lin = (vx, vy)
        return self._lin

    def _all(self):
        yield self.stem
        for b in self.blocks:
            yield from b

    # ------------------------------------------------------------------ forward
    def _forward(self, x, train: bool, want_grad: bool):
        with ops.precision(self.split):
            return self._forward_impl(x, train, want_grad)

    def _backward(self, saved, hm, pts, d_hm, d_pts, handle: Optional["_HeadHandle"], logits_grad=None,
                  do_allreduce=True, persistent_arena=False):
        with ops.precision(self.split):
            return self._backward_impl(saved, hm, pts, d_hm, d_pts, handle, logits_grad, do_allreduce,
                                       persistent_arena)

    def _forward_impl(self, x, train: bool, want_grad: bool):
        m = self.model
        dev = x.device
        self._setup(dev)
        self._packs.pack_all(want_grad)
        out_wpk, out_wpk_t = self._packs.wpk[id(m.out)], self._packs.wpk_t[id(m.out)]
        sc_ = self.stem.conv
        if ops.use_image_path(sc_.in_channels, sc_.kernel_size[0], 1, sc_.padding[0], sc_.dilation[0], sc_.out_channels):
            xin = x.contiguous().float()  # bf16 mode: the stem reads the image itself (no patch matrix)
        else:
            xin = ops.im2col_nchw(x, 7, 1, 3)  # explicit im2col + 1x1 conv over [.., 192] patches
        saved = {"x": xin, "blocks": [], "out_wpk_t": out_wpk_t}
        if train:
            y0 = self.stem.fwd_train(xin)
            a = ops.bn_apply_act(y0, self.stem.scale, self.stem.shift, ops.ACT_RELU, 0.0)
            saved["y0"] = y0
            for c1, c2, cs in self.blocks:
                y1 = c1.fwd_train(a)
                a1 = ops.bn_apply_act(y1, c1.scale, c1.shift, ops.ACT_RELU, 0.0)
                y2 = c2.fwd_train(a1)
                ys = cs.fwd_train(a)
                out = ops.bn_apply_act(ys, cs.scale, cs.shift, ops.ACT_RELU, 0.0, y2=y2, scale2=c2.scale,
                                       shift2=c2.shift)
                saved["blocks"].append((a, y1, a1, y2, ys, out))
                a = out
        else:
            sc, sh = self.stem.eval_affine()
            if xin.dtype == torch.float32:
                a = ops.conv_image_fwd(xin, self.stem.wpk, self.stem.cout, sc_.kernel_size[0], sc_.padding[0],
                                       scale=sc, shift=sh, act=ops.ACT_RELU)
            else:
                a = ops.conv_fwd(xin, self.stem.wpk, self.stem.cout, 1, 1, 0, scale=sc, shift=sh, act=ops.ACT_RELU)
            for c1, c2, cs in self.blocks:
                sc, sh = c1.eval_affine()
                a1 = ops.conv_fwd(a, c1.wpk, c1.cout, 3, 1, 2, 2, scale=sc, shift=sh, act=ops.ACT_RELU)
                sc, sh = cs.eval_affine()
                t = ops.conv_fwd(a, cs.wpk, cs.cout, 1, 1, 0, scale=sc, shift=sh)
                sc, sh = c2.eval_affine()
                a = ops.conv_fwd(a1, c2.wpk, c2.cout, 3, 1, 1, scale=sc, shift=sh, residual=t, act=ops.ACT_RELU)
        saved["a_last"] = a
        k = m.out.out_channels
        logits = ops.conv_fwd(a, out_wpk, k, 1, 1, 0, out_dtype=torch.float32, shift=m.out.bias.detach(), nchw_out=True)
        return logits, saved

    def _softmax(self, logits):
        b, k, h, w = logits.shape
        vx, vy = self._coords(logits.device, h, w)
        hm = torch.empty_like(logits)
        pts = torch.empty(b, k, 2, dtype=torch.float32, device=logits.device)
        lib().call("b200cv_kpt_softmax_argmax", ptr(logits), ptr(vx), ptr(vy), ptr(hm), ptr(pts), b * k, h, w,
                   stream_ptr())
        return hm, pts

    # ------------------------------------------------------------------ backward
    def _backward_impl(self, saved, hm, pts, d_hm, d_pts, handle: Optional[_HeadHandle], logits_grad=None,
                       do_allreduce=True, persistent_arena=False):
        m = self.model
        dev = hm.device if hm is not None else logits_grad.device
        arena = self._arena
        if not persistent_arena and arena.aliased_by_param_grads():
            arena = GradArena(self.params, dev)
            packs = ConvPackSet(self._conv_list(), dev, arena, flat=[self.stem.conv], split=self.split)
        else:
            packs = self._packs
        packs.zero_grads()
        side = None
        if not self.split and os.environ.get("B200CV_WGRAD_TAIL_FILL", "52") != "0":
            if getattr(self, "_side", None) is None or self._side.side.device != dev:
                self._side = _SideQueue(dev)
            side = self._side
            side.begin()
        views, gview = arena.views, arena.view_of
        a_last = saved["a_last"]
        b, h, w = a_last.shape[0], a_last.shape[1], a_last.shape[2]
        k = m.out.out_channels
        if logits_grad is not None:  # onnx_mode: gradient of the raw logits (NCHW fp32) given directly
            dl = ops.nchw_to_nhwc(logits_grad, 16)
        else:
            vx, vy = self._coords(dev, h, w)
            dl_ld = 16 * ops.split_pieces() if self.split else 16
            dl = torch.empty(b, h, w, dl_ld, dtype=torch.bfloat16, device=dev)
            p = handle.pending if handle is not None else None
            if p is not None:
                handle.pending = None
                lib().call("b200cv_kpt_head_bwd", ptr(hm), ptr(p["thm"]), ptr(pts), ptr(p["tpts"]), ptr(p["ubar"]),
                           ptr(vx), ptr(vy), ptr(p["g_loc"]), ptr(p["g_geo"]), ptr(d_hm), ptr(d_pts), b, k, h, w,
                           p["loss_type"], int(p["include_geo"]), float(p["gamma_h"]), float(p["gamma_v"]), ptr(dl),
                           dl_ld, stream_ptr())
            else:
                zeros = torch.zeros(b, k, 2, dtype=torch.float32, device=dev)
                lib().call("b200cv_kpt_head_bwd", ptr(hm), None, ptr(pts), ptr(zeros), None, ptr(vx), ptr(vy), None,
                           None, ptr(d_hm), ptr(d_pts), b, k, h, w, 0, 0, 0.0, 0.0, ptr(dl), dl_ld, stream_ptr())
        # head conv (bias, linear)
        gview[id(m.out.bias)].copy_(ops.bias_grad(dl, k))
        ops.conv_wgrad(a_last, dl, k, 1, 1, 0, out=packs.dwp[id(m.out)])
        g = ops.conv_dgrad(dl, saved["out_wpk_t"], m.out.in_channels, 1, 1, 0, 1, (h, w))
        fuse = os.environ.get("B200CV_FUSE_BN_REDUCE", "2") != "0" and not self.split
        stem_parts = None
        for bi, ((c1, c2, cs), (a_in, y1, a1, y2, ys, out)) in enumerate(
                zip(reversed(self.blocks), reversed(saved["blocks"]))):
            # out = relu(bn_s(ys) + bn_2(y2)): both branches see dz = g * relu'(out)
            dual = os.environ.get("B200CV_DUAL_BN", "1") != "0"
            if dual:
                # bn_s and bn_2 share da = g and the ReLU mask of `out`: one reduction pass and one apply pass for both
                count = ys.numel() // ys.shape[-1]
                parts_s, parts_2 = ops.bn_bwd_reduce2(g, out, ys, y2, cs.mean, cs.rstd, c2.mean, c2.rstd,
                                                      ops.ACT_RELU, 0.0)
                cs.finalize_bwd(parts_s, count, gview)
                c2.finalize_bwd(parts_2, count, gview)
                dys, dy2 = ops.bn_bwd_apply2(g, out, ys, y2, cs.mean, cs.rstd, c2.mean, c2.rstd, cs.coef, c2.coef,
                                             ops.ACT_RELU, 0.0)
                g_in = cs.finish(a_in, dys, gview, packs, side=side)
            else:
                g_in = cs.bwd(a_in, ys, g, out, ops.ACT_RELU, gview, packs, side=side)
            # a1 = relu(bn_1(y1)) has ONE consumer: conv2's data gradient is dL/da1, so its epilogue also forms the
            # BN-backward sums of bn_1; likewise the last data gradient into the first block's input for the stem
            red1, parts1 = c1.reduce_spec(y1, ops.ACT_RELU) if fuse else (None, None)
            if dual:
                g_a1 = c2.finish(a1, dy2, gview, packs, bn_reduce=red1, side=side)
            else:
                g_a1 = c2.bwd(a1, y2, g, out, ops.ACT_RELU, gview, packs, bn_reduce=red1, side=side)
            red0 = None
            if fuse and bi == len(self.blocks) - 1:
                red0, stem_parts = self.stem.reduce_spec(saved["y0"], ops.ACT_RELU)
            g = c1.bwd(a_in, y1, g_a1, None, ops.ACT_RELU, gview, packs, dx_out=g_in, parts=parts1, bn_reduce=red0,
                       side=side)
        self.stem.bwd(saved["x"], saved["y0"], g, None, ops.ACT_RELU, gview, packs, want_dx=False, parts=stem_parts)
        if side is not None:
            side.join()
        packs.unpack_all()
        if do_allreduce:
            allreduce_gradients(arena.flat)
        return views

    # ------------------------------------------------------------------ public entry point
    def run(self, x):
        require_cuda(x, "KeypointNet.forward")
        m = self.model
        with torch.cuda.device(x.device):  # launches go to the tensors' device, whatever the current one is
            if m.onnx_mode:
                return _KeypointLogitsFn.apply(self, x.float(), m.training, torch.is_grad_enabled(), *self.params)
            handle = _HeadHandle()
            step = self._graphed_step(x)
            if step is not None:
                hm, pts = _KeypointGraphFn.apply(step, x, handle, *self.params)
            else:
                hm, pts = _KeypointNetFn.apply(self, x.float(), m.training, handle, torch.is_grad_enabled(),
                                               *self.params)
        hm._b200cv_head = handle
        pts._b200cv_head = handle
        return hm, pts

    def _graphed_step(self, x):
        """The CUDA-graph step for this input shape, once two eager training steps have run with it (None = take the
        eager launches).  ~100 launches per step: at the reference's batch sizes (8-32) the host cannot keep up."""
        if not (self.model.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.params)):
            return None
        if os.environ.get("B200CV_CUDA_GRAPH", "1") == "0" or torch.cuda.is_current_stream_capturing():
            return None
        graphs = self.__dict__.setdefault("_graphs", {})
        key = (tuple(x.shape), str(x.device), self.split)
        entry = graphs.get(key, 0)
        if isinstance(entry, _GraphedRektStep):
            return entry if entry.usable() else None
        if entry >= GRAPH_WARMUP_CALLS:
            try:
                graphs[key] = _GraphedRektStep(self, x)
                return graphs[key]
            except Exception as e:  # capture is an optimisation: fall back to the eager launches, loudly
                import warnings

                warnings.warn(f"b200cv: CUDA-graph capture of the KeypointNet step failed ({e}); staying eager")
                graphs[key] = -(10 ** 9)
                torch.cuda.synchronize()
                return None
        graphs[key] = entry + 1
        return None


GRAPH_WARMUP_CALLS = 2


class _GraphedRektStep:
    """One KeypointNet training step (fixed input shape) as CUDA graphs sharing a memory pool: the forward graph
    (network + soft-argmax) is captured at the third step with a shape; the backward graph -- loss gradient, fused head
    backward, the whole backbone -- at the first backward whose CrossRatioLoss configuration is known, one per
    configuration.  The heat-maps / points handed to the caller are views of the step's static buffers: valid until
    the next training forward with the same shape (every training loop consumes them at once)."""

    def __init__(self, engine, x):
        self.engine = engine
        dev = x.device
        self.static_x = x.detach().float().clone()
        engine._setup(dev)
        self.weight_ptrs = tuple(p.data_ptr() for p in engine.params)
        torch.cuda.synchronize()
        self.pool = torch.cuda.graph_pool_handle()
        self.fwd_graph = torch.cuda.CUDAGraph()
        n0 = lib().launches
        with torch.no_grad():
            with torch.cuda.graph(self.fwd_graph, pool=self.pool):
                logits, self.saved = engine._forward(self.static_x, True, True)
                self.hm, self.pts = engine._softmax(logits)
        self.fwd_launches = lib().launches - n0
        self.bwd = {}  # loss configuration -> (graph, static loss tensors, views, launches)
        self.generation = 0
        torch.cuda.synchronize()

    def usable(self) -> bool:
        e = self.engine
        return (tuple(p.data_ptr() for p in e.params) == self.weight_ptrs and e._arena is not None
                and not e._arena.aliased_by_param_grads())

    def backward(self, handle, d_hm, d_pts):
        """Replay (or first capture) the backward graph for the pending loss; None = not graphable (the caller runs the
        eager backward on the saved static activations)."""
        p = handle.pending if handle is not None else None
        if p is None or d_hm is not None or not self.usable():
            return None
        key = (p["loss_type"], bool(p["include_geo"]), float(p["gamma_h"]), float(p["gamma_v"]), p["thm"] is None)
        names = ("thm", "tpts", "ubar", "g_loc", "g_geo")
        entry = self.bwd.get(key)
        if entry is None:
            static = {k: (p[k].detach().clone() if p[k] is not None else None) for k in names}
            static_dpts = torch.zeros_like(self.pts) if d_pts is None else d_pts.detach().clone()
            pend = dict(p)
            pend.update(static)
            h = _HeadHandle()
            h.pending = pend
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            n0 = lib().launches
            with torch.no_grad():
                with torch.cuda.graph(g, pool=self.pool):
                    views = self.engine._backward(self.saved, self.hm, self.pts, None, static_dpts, h,
                                                  do_allreduce=False, persistent_arena=True)
            entry = (g, static, static_dpts, views, lib().launches - n0)
            self.bwd[key] = entry
        g, static, static_dpts, views, launches = entry
        for k in names:
            if static[k] is not None:
                static[k].copy_(p[k], non_blocking=True)
        if d_pts is None:
            static_dpts.zero_()
        else:
            static_dpts.copy_(d_pts, non_blocking=True)
        handle.pending = None
        g.replay()
        lib().launches += launches
        allreduce_gradients(self.engine._arena.flat)
        return views


class _KeypointGraphFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, step, x, handle, *params):
        step.static_x.copy_(x, non_blocking=True)
        step.fwd_graph.replay()
        lib().launches += step.fwd_launches
        for c in step.engine._all():  # the replay moved the running statistics: drop the folded inference affines
            c._eval_key = None
        step.generation += 1
        ctx.step, ctx.handle, ctx.generation = step, handle, step.generation
        ctx.set_materialize_grads(False)
        return step.hm.view_as(step.hm), step.pts.view_as(step.pts)

    @staticmethod
    def backward(ctx, d_hm, d_pts):
        step = ctx.step
        if ctx.generation != step.generation:
            raise RuntimeError("b200cv: this KeypointNet forward was replayed from a CUDA graph and a LATER forward of "
                               "the same shape has overwritten its saved activations; back-propagate each forward "
                               "before the next one, or set B200CV_CUDA_GRAPH=0")
        d_hm = d_hm.contiguous().float() if d_hm is not None else None
        d_pts = d_pts.contiguous().float() if d_pts is not None else None
        with torch.cuda.device(step.static_x.device):
            views = step.backward(ctx.handle, d_hm, d_pts)
            if views is None:  # un-fused loss / upstream heat-map gradient: eager backward on the static activations
                views = step.engine._backward(step.saved, step.hm, step.pts, d_hm, d_pts, ctx.handle)
        return (None, None, None, *views)


class _KeypointNetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine, x, train, handle, grad_enabled, *params):
        # eval-mode BatchNorm (folded running statistics) has no backward here: such a pass is computed like the
        # reference computes it (RektNet/detect.py:38-39 calls it without no_grad) and returned as a constant
        want_grad = grad_enabled and train and any(ctx.needs_input_grad[5:])
        logits, saved = engine._forward(x, train, want_grad)
        hm, pts = engine._softmax(logits)
        if not want_grad:
            ctx.mark_non_differentiable(hm, pts)
            return hm, pts
        ctx.engine, ctx.saved, ctx.handle = engine, saved, handle
        ctx.save_for_backward(hm, pts)
        ctx.set_materialize_grads(False)  # an output the loss does not use arrives as None, not as a zero tensor
        return hm, pts

    @staticmethod
    def backward(ctx, d_hm, d_pts):
        hm, pts = ctx.saved_tensors
        d_hm = d_hm.contiguous().float() if d_hm is not None else None
        d_pts = d_pts.contiguous().float() if d_pts is not None else None
        with torch.cuda.device(hm.device):
            views = ctx.engine._backward(ctx.saved, hm, pts, d_hm, d_pts, ctx.handle)
        ctx.saved = None
        return (None, None, None, None, None, *views)


class _KeypointLogitsFn(torch.autograd.Function):
    """onnx_mode=True: forward returns the raw head logits (keypoint_net.py:65-66)."""

    @staticmethod
    def forward(ctx, engine, x, train, grad_enabled, *params):
        want_grad = grad_enabled and train and any(ctx.needs_input_grad[4:])
        logits, saved = engine._forward(x, train, want_grad)
        if not want_grad:
            ctx.mark_non_differentiable(logits)
            return logits
        ctx.engine, ctx.saved = engine, saved
        return logits

    @staticmethod
    def backward(ctx, d_logits):
        with torch.cuda.device(d_logits.device):
            views = ctx.engine._backward(ctx.saved, None, None, None, None, None,
                                         logits_grad=d_logits.contiguous().float())
        ctx.saved = None
        return (None, None, None, None, *views)


class ResBlockEngine:
    """A single ``resnet.ResNet`` block called on its own (RektNet/resnet.py:22-27): NCHW fp32 in, NCHW fp32 out,
    gradients for the block's parameters AND its input.  Same kernels and the same per-layer objects as inside
    RektNetEngine; the block boundary costs one layout conversion each way."""

    def __init__(self, block):
        self.block = block
        self.c1 = _ConvBN(block.conv1, block.bn1)
        self.c2 = _ConvBN(block.conv2, block.bn2)
        self.cs = _ConvBN(block.shortcut_conv, block.shortcut_bn)
        if ops.pad_channels(self.c1.cout) != self.c1.cout:
            raise ValueError(f"ResNet block with {self.c1.cout} output channels: the B200 layout needs 16, 32 or a "
                             "multiple of 64 channels in every normalised layer")
        self.params = list(block.parameters())
        self.split = ops.default_split()
        self._arena = None
        self._packs = None

    def _setup(self, dev):
        if self._arena is None or self._arena.flat.device != dev:
            self._arena = GradArena(self.params, dev)
            convs = [(c.conv, True) for c in (self.c1, self.c2, self.cs)]
            self._packs = ConvPackSet(convs, dev, self._arena, split=self.split)
            for c in (self.c1, self.c2, self.cs):
                c.wpk, c.wpk_t = self._packs.wpk[id(c.conv)], self._packs.wpk_t[id(c.conv)]

    def forward(self, x, train: bool, want_grad: bool):
        self._setup(x.device)
        self._packs.pack_all(want_grad)
        c1, c2, cs = self.c1, self.c2, self.cs
        a = ops.nchw_to_nhwc(x)
        if train:
            y1 = c1.fwd_train(a)
            a1 = ops.bn_apply_act(y1, c1.scale, c1.shift, ops.ACT_RELU, 0.0)
            y2 = c2.fwd_train(a1)
            ys = cs.fwd_train(a)
            out = ops.bn_apply_act(ys, cs.scale, cs.shift, ops.ACT_RELU, 0.0, y2=y2, scale2=c2.scale, shift2=c2.shift)
            saved = (a, y1, a1, y2, ys, out)
        else:
            sc, sh = c1.eval_affine()
            a1 = ops.conv_fwd(a, c1.wpk, c1.cout, 3, 1, 2, 2, scale=sc, shift=sh, act=ops.ACT_RELU)
            sc, sh = cs.eval_affine()
            t = ops.conv_fwd(a, cs.wpk, cs.cout, 1, 1, 0, scale=sc, shift=sh)
            sc, sh = c2.eval_affine()
            out = ops.conv_fwd(a1, c2.wpk, c2.cout, 3, 1, 1, scale=sc, shift=sh, residual=t, act=ops.ACT_RELU)
            saved = None
        return ops.nhwc_to_nchw(out, c1.cout), saved

    def backward(self, saved, d_out):
        a_in, y1, a1, y2, ys, out = saved
        c1, c2, cs = self.c1, self.c2, self.cs
        dev = d_out.device
        arena, packs = self._arena, self._packs
        if arena.aliased_by_param_grads():
            arena = GradArena(self.params, dev)
            packs = ConvPackSet([(c.conv, True) for c in (c1, c2, cs)], dev, arena, split=self.split)
        packs.zero_grads()
        gview = arena.view_of
        g = ops.nchw_to_nhwc(d_out, ops.channels(out))
        count = ys.numel() // ys.shape[-1]
        parts_s, parts_2 = ops.bn_bwd_reduce2(g, out, ys, y2, cs.mean, cs.rstd, c2.mean, c2.rstd, ops.ACT_RELU, 0.0)
        cs.finalize_bwd(parts_s, count, gview)
        c2.finalize_bwd(parts_2, count, gview)
        dys, dy2 = ops.bn_bwd_apply2(g, out, ys, y2, cs.mean, cs.rstd, c2.mean, c2.rstd, cs.coef, c2.coef,
                                     ops.ACT_RELU, 0.0)
        g_in = cs.finish(a_in, dys, gview, packs)
        g_a1 = c2.finish(a1, dy2, gview, packs)
        g_in = c1.bwd(a_in, y1, g_a1, None, ops.ACT_RELU, gview, packs, dx_out=g_in)
        packs.unpack_all()
        allreduce_gradients(arena.flat)
        return ops.nhwc_to_nchw(g_in, c1.cin), arena.views

    def run(self, x):
        require_cuda(x, "ResNet.forward")
        with torch.cuda.device(x.device):
            return _ResBlockFn.apply(self, x.float(), self.block.training, torch.is_grad_enabled(), *self.params)


class _ResBlockFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine, x, train, grad_enabled, *params):
        want_grad = grad_enabled and train and any(ctx.needs_input_grad[1:])
        with ops.precision(engine.split):
            out, saved = engine.forward(x, train, want_grad)
        if not want_grad:
            ctx.mark_non_differentiable(out)
            return out
        ctx.engine, ctx.saved = engine, saved
        return out

    @staticmethod
    def backward(ctx, d_out):
        with torch.cuda.device(d_out.device), ops.precision(ctx.engine.split):
            dx, views = ctx.engine.backward(ctx.saved, d_out.contiguous().float())
        ctx.saved = None
        return (None, dx, None, None, *views)


# ---------------------------------------------------------------------------- CrossRatioLoss
def _loss_forward(hm, pts, thm, tpts, loss_type, include_geo, gamma_h, gamma_v):
    b, k = pts.shape[0], pts.shape[1]
    hw = hm.shape[-1] * hm.shape[-2] if hm is not None else 0
    dev = pts.device
    loss3 = torch.empty(3, dtype=torch.float32, device=dev)
    ubar = torch.zeros(18, dtype=torch.float32, device=dev)
    ws = torch.zeros(1, dtype=torch.float64, device=dev)
    lib().call("b200cv_kpt_loss", ptr(hm), ptr(pts), ptr(thm), ptr(tpts), b, k, hw, loss_type, int(include_geo),
               float(gamma_h), float(gamma_v), ptr(ws), ptr(loss3), ptr(ubar), stream_ptr())
    return loss3, ubar


class CrossRatioLossFn(torch.autograd.Function):
    """(location, geo, total) = f(hm, pts); the backward is fused into KeypointNet's when possible."""

    @staticmethod
    def forward(ctx, hm, pts, thm, tpts, loss_type, include_geo, gamma_h, gamma_v, handle):
        require_cuda(pts, "CrossRatioLoss.forward")
        hm_c, pts_c = hm.contiguous().float(), pts.contiguous().float()
        thm_c = thm.contiguous().float() if thm is not None else None
        tpts_c = tpts.contiguous().float()
        loss3, ubar = _loss_forward(hm_c, pts_c, thm_c, tpts_c, loss_type, include_geo, gamma_h, gamma_v)
        ctx.save_for_backward(hm_c, pts_c, thm_c, tpts_c, ubar)
        ctx.cfg = (loss_type, include_geo, gamma_h, gamma_v)
        ctx.handle = handle
        return loss3

    @staticmethod
    def backward(ctx, g3):
        hm, pts, thm, tpts, ubar = ctx.saved_tensors
        loss_type, include_geo, gamma_h, gamma_v = ctx.cfg
        g3 = g3.contiguous().float()
        g_loc = (g3[0:1] + g3[2:3]).contiguous()  # total = location + geo
        g_geo = (g3[1:2] + g3[2:3]).contiguous()
        b, k = pts.shape[0], pts.shape[1]
        if ctx.handle is not None and ctx.handle.pending is None:
            # defer: KeypointNet's backward runs the fused kernel.  A zero d_pts keeps autograd flowing.
            ctx.handle.pending = dict(thm=thm, tpts=tpts, ubar=ubar, g_loc=g_loc, g_geo=g_geo, loss_type=loss_type,
                                      include_geo=include_geo, gamma_h=gamma_h, gamma_v=gamma_v)
            return (None, torch.zeros_like(pts), None, None, None, None, None, None, None)
        d_pts = torch.empty_like(pts)
        d_hm = torch.empty_like(hm) if (loss_type == 1 and ctx.needs_input_grad[0]) else None
        lib().call("b200cv_kpt_loss_bwd", ptr(hm), ptr(thm), ptr(pts), ptr(tpts), ptr(ubar), ptr(g_loc), ptr(g_geo), b,
                   k, hm.shape[-2], hm.shape[-1], loss_type, int(include_geo), float(gamma_h), float(gamma_v),
                   ptr(d_pts), ptr(d_hm), stream_ptr())
        return (d_hm, d_pts, None, None, None, None, None, None, None)
