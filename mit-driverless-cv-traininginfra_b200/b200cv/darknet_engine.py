"""Executor of a Darknet cfg on the B200 kernels.

``models.Darknet`` keeps ordinary fp32 ``nn.Conv2d`` / ``nn.BatchNorm2d`` modules (so state_dict,
``.weights`` I/O and torch.optim work unchanged) but never calls them: forward and backward of the
whole network are ONE ``torch.autograd.Function`` whose body is a sequence of C-ABI launches over
NHWC bf16 activations:

  conv (tcgen05 implicit GEMM, epilogue emits per-channel sum/sumsq) -> bn_finalize ->
  bn_apply + LeakyReLU (+ fused shortcut add) ... -> head conv (fp32 logits) -> yolo targets + loss

and, in reverse, yolo dlogits -> [bn_bwd_reduce, bn_bwd_finalize, bn_bwd_apply, wgrad, dgrad(+fan-in
residual)] per conv.  Parameter gradients land in one flat fp32 arena (views are returned to
autograd), which is what the data-parallel all-reduce sends over NCCL.
"""
from __future__ import annotations

import weakref
from typing import List, Optional

import torch

from . import ops, yolo_ops
from .lib import require_cuda
from .packing import ConvPackSet, GradArena
from . import parallel
from .parallel import allreduce_gradients

BN_EPS = 1e-5
BN_MOMENTUM = 0.1
GRAPH_WARMUP_CALLS = 2  # eager steps with a given input shape before the step is captured as CUDA graphs


def _fuse_bn_reduce() -> int:
    """0: separate bn_bwd_reduce passes; 1: fused into the 3x3 data-gradient epilogues; 2: into every stride-1 one."""
    import os

    return int(os.environ.get("B200CV_FUSE_BN_REDUCE", "2"))


def _wgrad_side_stream() -> bool:
    import os

    # opt-in: measured no gain on B200 (26.5 vs 26.6 ms/step) -- a wgrad CTA (~200 KB of shared memory) cannot share
    # an SM with the persistent dgrad CTAs, and the BN passes it could overlap are short
    return os.environ.get("B200CV_WGRAD_STREAM", "0") != "0"


def _wgrad_tail_fill_hw() -> int:
    """Layers whose feature map is at most this many pixels high launch their weight gradient on the side stream AFTER
    their data gradient: the persistent dgrad of a 13x13 layer is 170 tiles on 148 SMs, so 126 SMs idle through its
    second wave -- the wgrad CTAs queued behind it fill them, and they keep the tensor pipe busy under the HBM-bound
    BatchNorm-backward passes of the next layer.  Measured on Darknet-53 416^2 bs64: 25.77 ms/step off, 25.75 / 25.61 /
    25.20 ms with 13 / 26 / 52, no further gain at 104+ (those weight gradients are HBM-bound themselves; launching
    wgrad BEFORE dgrad, B200CV_WGRAD_STREAM=1, was neutral).  0 = off."""
    import os

    return int(os.environ.get("B200CV_WGRAD_TAIL_FILL", "52"))


def _graphs_enabled() -> bool:
    import os

    return os.environ.get("B200CV_CUDA_GRAPH", "1") != "0"


class _Layer:
    __slots__ = ("index", "type", "conv", "bn", "act", "slope", "k", "stride", "pad", "cin", "cout", "inputs",
                 "post_from", "fused_alias", "yolo", "stats", "bstats", "scale", "shift", "mean", "rstd", "sums",
                 "coef", "wpk", "wpk_t", "pool_stride", "eval_key", "eval_affine")

    def __init__(self, index, type_):
        self.index, self.type = index, type_
        self.conv = self.bn = self.yolo = None
        self.post_from = None      # conv: index of the tensor added after the activation (fused shortcut)
        self.fused_alias = False   # shortcut: output is the previous conv's post-add output
        self.inputs: List[int] = []
        self.stats = None
        self.eval_key = self.eval_affine = None


class DarknetEngine:
    def __init__(self, model):
        self.model = model
        self.layers: List[_Layer] = []
        slope = float(model.hyperparams["leaky_slope"])
        n = len(model.module_defs)
        for i, (d, m) in enumerate(zip(model.module_defs, model.module_list)):
            L = _Layer(i, d["type"])
            if L.type == "convolutional":
                L.conv = m[0]
                L.bn = m[1] if d["filters"] != "preyolo" else None
                names = [type(x).__name__ for x in m]
                L.act = ops.ACT_LEAKY if "LeakyReLU" in names else (ops.ACT_RELU if "ReLU" in names else ops.ACT_NONE)
                L.slope = slope if L.act == ops.ACT_LEAKY else 0.0
                L.k, L.stride, L.pad = L.conv.kernel_size[0], L.conv.stride[0], L.conv.padding[0]
                L.cin, L.cout = L.conv.in_channels, L.conv.out_channels
                L.inputs = [i - 1]
            elif L.type == "maxpool":
                L.pool_stride = int(d["stride"])
                if int(d["size"]) != 2 or L.pool_stride not in (1, 2):
                    raise ValueError("only 2x2 max-pool with stride 1 or 2 is supported")
                L.inputs = [i - 1]
            elif L.type == "upsample":
                if int(d["stride"]) != 2:
                    raise ValueError("only x2 upsample is supported")
                L.inputs = [i - 1]
            elif L.type == "route":
                L.inputs = [(i + v) if v < 0 else v for v in (int(x) for x in d["layers"].split(","))]
            elif L.type == "shortcut":
                f = int(d["from"])
                L.inputs = [i - 1, (i + f) if f < 0 else f]
            elif L.type == "yolo":
                L.yolo = m[0]
                L.inputs = [i - 1]
            else:
                raise ValueError(f"unsupported cfg block [{L.type}]")
            self.layers.append(L)
        # fuse "conv+bn+act ; shortcut" into the conv's apply pass when nothing else reads the conv output
        consumers = {i: [] for i in range(-1, n)}
        for L in self.layers:
            for j in L.inputs:
                consumers[j].append(L.index)
        for L in self.layers:
            if L.type == "shortcut":
                prev = self.layers[L.index - 1]
                if prev.type == "convolutional" and prev.bn is not None and consumers[prev.index] == [L.index]:
                    prev.post_from = L.inputs[1]
                    L.fused_alias = True
        self.params = list(model.parameters())
        # "fp32" = the fp32-parity mode (split bf16x3 operands, see ops.py); chosen when the engine is created
        # (B200CV_PRECISION / ops.set_default_precision) or later through set_precision()
        self.split = ops.default_split()
        self._arena = None
        self._packs = None
        self._anchor_cache = {}
        self._pack_key = None   # versions/pointers of the conv weights at the last inference-pass pack
        self._graphs = {}       # (shapes) -> _GraphedStep | int (eager warm-up calls seen so far)

    # ------------------------------------------------------------------ helpers
    @property
    def precision(self) -> str:
        return "fp32" if self.split else "bf16"

    def set_precision(self, name: str):
        """"bf16" (default) or "fp32" (bf16x3 split operands: parity with the reference's fp32 arithmetic)."""
        if name not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        if (name == "fp32") != self.split:
            self.split = name == "fp32"
            self._arena = self._packs = self._pack_key = None  # operand layouts change: rebuild packs and graphs
            self._graphs = {}
        return self

    def _setup(self, dev):
        if self._arena is None or self._arena.flat.device != dev:
            # per-channel vectors of every BN layer; the partial-statistics matrices (forward: sum / sum of squares
            # from the conv epilogue, backward: sum dz / sum dz*xhat) live in two flat arenas so that one memset
            # per pass clears all of them
            bn_layers = [L for L in self.layers if L.type == "convolutional" and L.bn is not None]
            for L in bn_layers:
                if ops.pad_channels(L.cout) != L.cout:
                    raise ValueError(f"BatchNorm over {L.cout} channels: the B200 layout needs 16, 32 or a multiple "
                                     "of 64 channels in every normalised layer")
            per = [ops.STAT_PARTS * 2 * L.cout * ops.STAT_WORDS for L in bn_layers]
            self._fstat_arena = torch.zeros(sum(per), dtype=torch.int64, device=dev)
            self._bstat_arena = torch.zeros(sum(per), dtype=torch.int64, device=dev)
            off = 0
            f = lambda n: torch.empty(n, dtype=torch.float32, device=dev)
            for L, n in zip(bn_layers, per):
                c = L.cout
                L.stats = self._fstat_arena[off:off + n].view(ops.STAT_PARTS, 2 * c, ops.STAT_WORDS)
                L.bstats = self._bstat_arena[off:off + n].view(ops.STAT_PARTS, 2 * c, ops.STAT_WORDS)
                L.scale, L.shift, L.mean, L.rstd, L.coef = f(c), f(c), f(c), f(c), f(3 * c)
                off += n
            self._nbt = [L.bn.num_batches_tracked for L in bn_layers if L.bn.num_batches_tracked is not None]
            self._arena = GradArena(self.params, dev)
            convs = [(L.conv, L.index > 0) for L in self.layers if L.type == "convolutional"]
            self._packs = ConvPackSet(convs, dev, self._arena, flat=self._flat_convs(), split=self.split, d2s=self._d2s_convs())
            for L in self.layers:
                if L.type == "convolutional":
                    L.wpk, L.wpk_t = self._packs.wpk[id(L.conv)], self._packs.wpk_t[id(L.conv)]

    def _bn_owner(self, j: int):
        """Index of the conv+BN layer whose activation output IS tensor outs[j] (so that the gradient written into
        grads[j] is that layer's dL/da), or None."""
        if j < 0:
            return None
        Lj = self.layers[j]
        if Lj.type == "convolutional" and Lj.bn is not None and Lj.post_from is None:
            return j
        if Lj.type == "shortcut" and Lj.fused_alias:
            return j - 1
        return None

    def _d2s_convs(self):
        """3x3 stride-2 convs whose data gradient can run as one depth-to-space launch (ops.conv_dgrad_d2s)."""
        return [L.conv for L in self.layers
                if L.type == "convolutional" and L.index > 0 and ops.d2s_dgrad_ok(L.cin, L.k, L.stride, L.pad, 1 << 30)]

    def _flat_convs(self):
        L0 = self.layers[0]
        return [L0.conv] if (L0.type == "convolutional" and ops.use_flat_path(L0.cin, L0.k)) else []

    def _pack(self, need_t: bool, frozen: bool = False):
        """Re-pack the fp32 weights into the bf16 operand layouts.  `frozen` (inference passes): skip the launch when
        no weight tensor was modified or re-homed since the last pack (tensor version counters + data pointers)."""
        if frozen:
            key = tuple((L.conv.weight._version, L.conv.weight.data_ptr()) for L in self.layers
                        if L.type == "convolutional")
            if key == self._pack_key:
                return
            self._packs.pack_all(need_t)
            self._pack_key = key
            return
        self._pack_key = None  # a training pass: the optimizer is about to change the weights
        self._packs.pack_all(need_t)

    def _eval_affine(self, L):
        """Inference BatchNorm folded into the conv epilogue: y*scale + shift with scale = gamma/sqrt(var+eps),
        shift = beta - mean*scale (models.py:64 in eval mode).  Cached per layer until gamma / beta / the running
        statistics change (version counters), so a steady-state inference pass launches no per-layer fold kernels."""
        bn = L.bn
        key = (bn.weight._version, bn.bias._version, bn.running_mean._version, bn.running_var._version,
               bn.weight.data_ptr(), bn.running_mean.data_ptr())
        if L.eval_key != key:
            scale = bn.weight.detach() * torch.rsqrt(bn.running_var + BN_EPS)
            L.eval_affine = (scale, bn.bias.detach() - bn.running_mean * scale)
            L.eval_key = key
        return L.eval_affine

    def _anchors(self, L, gh, dev):
        key = (L.index, gh, str(dev))
        if key not in self._anchor_cache:
            self._anchor_cache[key] = yolo_ops.scaled_anchors(L.yolo.anchors, L.yolo.image_height / gh, dev)
        return self._anchor_cache[key]

    def _check_input(self, x):
        require_cuda(x, "Darknet.forward")
        if x.dim() != 4 or x.shape[1] != self.layers[0].cin:
            raise ValueError(f"expected input [B,{self.layers[0].cin},H,W], got {tuple(x.shape)}")

    # ------------------------------------------------------------------ forward
    def _run_forward(self, x, targets, bn_train: bool, want_grad: bool, private: bool = False):
        with ops.precision(self.split):
            return self._run_forward_impl(x, targets, bn_train, want_grad, private)

    def _run_backward(self, state, g7, force_persistent_arena=False, do_allreduce=True):
        """Eager backward.  With several processes the gradient arena is all-reduced (SUM) in buckets: the bucket of
        the LAST layers is complete first and its transfer runs on the collective's stream while the earlier layers are
        still being back-propagated."""
        handles = []
        gen = self._run_backward_segments(state, g7, force_persistent_arena, parallel.grad_buckets() if do_allreduce else 1)
        while True:
            try:
                chunk = next(gen)
            except StopIteration as done:
                views = done.value
                break
            if do_allreduce:
                handles.append(parallel.allreduce_gradients_async(chunk))
        parallel.finish_allreduces(handles)
        return views

    def _run_backward_segments(self, state, g7, force_persistent_arena, nbuckets):
        """Generator form of the backward pass: yields the slice of the flat gradient arena that has just become final
        (all its layers back-propagated, weight gradients un-packed), `nbuckets` times, last layers first; returns the
        per-parameter views.  The CUDA-graph step captures every segment as a graph of its own."""
        with ops.precision(self.split):
            gen = self._run_backward_impl(state, g7, force_persistent_arena, nbuckets)
            while True:
                try:
                    chunk = next(gen)
                except StopIteration as done:
                    return done.value
                _tls_leave = ops.split_mode()  # the consumer runs outside the precision context
                ops._tls.split = False
                try:
                    yield chunk
                finally:
                    ops._tls.split = _tls_leave

    def _bucket_plan(self, nbuckets):
        """[(first layer index, arena lo, arena hi, conv lo, conv hi)] of the gradient buckets in BACKWARD order, cut at
        layer boundaries.  Shares of the parameter bytes shrink geometrically (60 %, 24 %, 9.6 %, ...): most parameters
        sit in the LAST layers, whose backward is over after a fraction of the pass, while the early high-resolution
        layers take most of the time and own few parameters -- the bucket that is still exposed after backward is
        the smallest one."""
        layer_params, conv_index = [], []
        nconv = 0
        for L in self.layers:
            n = 0
            if L.type == "convolutional":
                n = sum(p.numel() for p in L.conv.parameters()) + (sum(p.numel() for p in L.bn.parameters())
                                                                  if L.bn is not None else 0)
                nconv += 1
            layer_params.append(n)
            conv_index.append(nconv)  # convs in layers [0, i]
        total = sum(layer_params)
        offs = [0]
        for n in layer_params:
            offs.append(offs[-1] + n)
        plan, hi_layer = [], len(self.layers)
        for k in range(nbuckets - 1):
            target = total * 0.4 ** (k + 1)  # arena offset where this bucket should start
            lo_layer = min(range(hi_layer), key=lambda i: abs(offs[i] - target)) if hi_layer > 0 else 0
            if lo_layer >= hi_layer or lo_layer <= 0:
                continue
            plan.append((lo_layer, offs[lo_layer], offs[hi_layer], conv_index[lo_layer - 1], conv_index[hi_layer - 1]))
            hi_layer = lo_layer
        plan.append((0, 0, offs[hi_layer], 0, conv_index[hi_layer - 1] if hi_layer > 0 else 0))
        return plan

    def _run_forward_impl(self, x, targets, bn_train: bool, want_grad: bool, private: bool = False):
        """Returns (out7 or detections, saved-state).  `private`: the per-layer BatchNorm vectors this pass saves for
        its backward are fresh tensors instead of the engine's static ones (a second forward issued while an earlier
        one still awaits its backward must not overwrite what that one saved)."""
        model = self.model
        dev = x.device
        self._setup(dev)
        self._pack(need_t=want_grad, frozen=not bn_train and not want_grad)
        flat0 = bool(self._flat_convs())
        L0 = self.layers[0]
        # bf16 mode: conv_0 reads the NCHW fp32 image itself (no patch matrix, csrc/conv_image.cu)
        img0 = flat0 and L0.bn is not None and L0.post_from is None and \
            ops.use_image_path(L0.cin, L0.k, L0.stride, L0.pad, 1, L0.cout)
        if img0:
            cur = x.contiguous().float()
        else:
            cur = ops.im2col_nchw(x, L0.k, L0.stride, L0.pad) if flat0 else ops.nchw_to_nhwc(x)
        outs: List[Optional[torch.Tensor]] = [None] * len(self.layers)
        saved = {}
        training = targets is not None
        out7 = torch.zeros(7, dtype=torch.float32, device=dev) if training else None
        if bn_train:
            for L in self.layers:  # our kernels update the running statistics behind torch's version counters
                L.eval_key = None
            self._fstat_arena.zero_()
            if self._nbt:
                torch._foreach_add_(self._nbt, 1)
        dets = []  # eval: per-head detections, concatenated at the end
        consts = (model.xy_loss, model.wh_loss, model.object_loss, model.no_object_loss)
        for L in self.layers:
            i = L.index
            if L.type == "convolutional":
                xin = cur
                k_, st_, pd_ = (1, 1, 0) if (flat0 and i == 0) else (L.k, L.stride, L.pad)
                if L.bn is not None:
                    post = outs[L.post_from] if L.post_from is not None else None
                    if bn_train:
                        if img0 and i == 0:
                            y = ops.conv_image_fwd(xin, L.wpk, L.cout, L.k, L.pad, stats=L.stats)
                        else:
                            y = ops.conv_fwd(xin, L.wpk, L.cout, k_, st_, pd_, stats=L.stats)
                        count = y.numel() // y.shape[-1]
                        vec = (L.scale, L.shift, L.mean, L.rstd)
                        if private:
                            vec = tuple(torch.empty_like(v) for v in vec)
                        cur = ops.bn_stats_apply_act(L.stats, count, L.bn.weight, L.bn.bias, None, BN_EPS, BN_MOMENTUM,
                                                     L.bn.running_mean, L.bn.running_var, vec[0], vec[1], vec[2],
                                                     vec[3], y, L.act, L.slope, post=post)
                        saved[i] = (xin, y, vec)
                    else:
                        scale, shift = self._eval_affine(L)
                        if img0 and i == 0:
                            cur = ops.conv_image_fwd(xin, L.wpk, L.cout, L.k, L.pad, scale=scale, shift=shift,
                                                     act=L.act, slope=L.slope)
                        else:
                            cur = ops.conv_fwd(xin, L.wpk, L.cout, k_, st_, pd_, scale=scale, shift=shift,
                                               residual=post, act=L.act, slope=L.slope, res_after_act=True)
                else:  # pre-YOLO conv: bias, linear, fp32 logits
                    cur = ops.conv_fwd(xin, L.wpk, L.cout, k_, st_, pd_, out_dtype=torch.float32,
                                       shift=L.conv.bias.detach())
                    saved[i] = (xin,)
            elif L.type == "maxpool":
                saved[i] = (cur,)
                cur = ops.maxpool_fwd(cur, L.pool_stride)
            elif L.type == "upsample":
                cur = ops.upsample_fwd(cur)
            elif L.type == "route":
                if len(L.inputs) == 1:
                    cur = outs[L.inputs[0]]
                else:
                    cur = ops.concat_channels([outs[j] for j in L.inputs])
            elif L.type == "shortcut":
                if L.fused_alias:
                    cur = outs[i - 1]
                else:
                    cur = outs[L.inputs[0]].clone()
                    ops.copy_slice(outs[L.inputs[1]], cur, accumulate=True)
            elif L.type == "yolo":
                z = cur  # fp32 [B,G,G,Cpad]
                yl = L.yolo
                gh, gw = z.shape[1], z.shape[2]
                stride = yl.image_height / gh
                sa = self._anchors(L, gh, dev)
                if training:
                    yt = yolo_ops.yolo_targets(targets, sa, gh, gw, yl.ignore_thres)
                    sums = torch.zeros(6, dtype=torch.float64, device=dev)
                    yolo_ops.yolo_loss_cells(z, False, yt, yl.num_classes, consts, sums=sums)
                    yolo_ops.yolo_loss_finalize(sums, yt, consts, out7)
                    saved[i] = (z, yt)
                    cur = None
                else:
                    rows = yl.num_anchors * gh * gw
                    d = torch.empty(z.shape[0], rows, 5 + yl.num_classes, dtype=torch.float32, device=dev)
                    yolo_ops.yolo_decode(z, False, yl.num_anchors, yl.num_classes, sa, stride, d, 0)
                    dets.append(d)
                    cur = d
            outs[i] = cur
        if training:
            return out7, (outs, saved, private)
        return torch.cat(dets, 1), None

    # ------------------------------------------------------------------ backward
    def _run_backward_impl(self, state, g7, force_persistent_arena=False, nbuckets=1):
        outs, saved, private = state
        plan = self._bucket_plan(nbuckets)
        cut = {first: (lo, hi, clo, chi) for first, lo, hi, clo, chi in plan}
        model = self.model
        dev = g7.device
        g = g7[0:1].contiguous().float()
        arena = self._arena
        if not force_persistent_arena and (private or arena.aliased_by_param_grads()):
            # param.grad still aliases the arena (zero_grad(set_to_none=False)): use a private arena for this
            # backward so autograd's in-place accumulation stays correct
            arena = GradArena(self.params, dev)
            packs = ConvPackSet([(L.conv, L.index > 0) for L in self.layers if L.type == "convolutional"], dev, arena,
                                flat=self._flat_convs(), split=self.split, d2s=self._d2s_convs())
        else:
            packs = self._packs
        packs.zero_grads()
        self._bstat_arena.zero_()
        views, gview = arena.views, arena.view_of
        consts = (model.xy_loss, model.wh_loss, model.object_loss, model.no_object_loss)
        grads: List[Optional[torch.Tensor]] = [None] * len(self.layers)
        reduced = set()  # BN layers whose backward sums were already produced by a dgrad epilogue
        fuse = _fuse_bn_reduce()
        # Weight gradients run on a side stream: wgrad(i) only needs dy(i) and its result is not read before the
        # un-pack at the end, so it overlaps dgrad(i) and -- more importantly -- the HBM-bound BatchNorm backward
        # passes of layer i-1 (tensor-bound and bandwidth-bound kernels share the SMs).  The operands are kept alive
        # until the join (no allocator reuse while the side stream may still read them).
        main = torch.cuda.current_stream()
        side = None
        tail_hw = 0 if self.split else _wgrad_tail_fill_hw()
        if _wgrad_side_stream() or tail_hw > 0:
            if getattr(self, "_side", None) is None or self._side.device != dev:
                self._side = torch.cuda.Stream(device=dev)
            side = self._side
            side.wait_stream(main)
        keep = []

        def wgrad(x_, dy_, cout_, k_, st_, pd_, out_, on_side=True, ev=None):
            if side is None or not on_side:
                ops.conv_wgrad(x_, dy_, cout_, k_, st_, pd_, out=out_)
                return
            if ev is None:
                ev = torch.cuda.Event()
                ev.record(main)
            side.wait_event(ev)
            with torch.cuda.stream(side):
                ops.conv_wgrad(x_, dy_, cout_, k_, st_, pd_, out=out_)
            keep.append((x_, dy_))

        def add_grad(j, t):
            if j < 0:
                return
            if grads[j] is None:
                grads[j] = t
            else:
                ops.copy_slice(t, grads[j], accumulate=True)

        for L in reversed(self.layers):
            i = L.index
            if (i + 1) in cut:  # every layer > i is done: that bucket of the gradient arena is final
                if side is not None:
                    main.wait_stream(side)
                    keep.clear()
                lo, hi, clo, chi = cut[i + 1]
                packs.unpack_range(clo, chi)
                yield arena.flat[lo:hi]
            if L.type == "yolo":
                z, yt = saved[i]
                dl = yolo_ops.yolo_head_grad(z, yt, L.yolo.num_classes, consts, g)
                add_grad(i - 1, dl)
                continue
            G = grads[i]
            if G is None:
                if L.type == "convolutional":  # parameters that received no gradient
                    gview[id(L.conv.weight)].zero_()
                    if L.bn is not None:
                        gview[id(L.bn.weight)].zero_()
                        gview[id(L.bn.bias)].zero_()
                    else:
                        gview[id(L.conv.bias)].zero_()
                continue
            if L.type == "convolutional":
                if L.bn is not None:
                    xin, y, (scale, shift, mean, rstd) = saved[i]
                    count = y.numel() // y.shape[-1]
                    parts = L.bstats
                    if i not in reduced:
                        ops.bn_bwd_reduce(G, y, None, scale, shift, mean, rstd, L.act, L.slope, partials=parts)
                    if i == 0 and xin.dtype == torch.float32 and L.post_from is None:
                        # conv_0 on the image: no data gradient, so dy = BN'(G, y) has one reader -- the weight
                        # gradient forms it in registers and the 709 MB tensor is never written
                        ops.conv_image_wgrad_bn(xin, G, y, parts, count, L.bn.weight, L.coef, gview[id(L.bn.weight)],
                                                gview[id(L.bn.bias)], scale, shift, mean, rstd, L.act, L.slope,
                                                L.cout, L.k, L.pad, 1, packs.dwp[id(L.conv)])
                        grads[i] = None
                        continue
                    dy = ops.bn_bwd_stats_apply(parts, count, L.bn.weight, L.coef, gview[id(L.bn.weight)],
                                                gview[id(L.bn.bias)], G, y, scale, shift, mean, rstd, L.act, L.slope)
                    if L.post_from is not None:
                        add_grad(L.post_from, G)
                else:
                    (xin,) = saved[i]
                    dy = G
                    gview[id(L.conv.bias)].copy_(ops.bias_grad(dy, L.cout))
                # tail-fill mode: the weight gradient of a small-map layer is queued BEHIND its data gradient
                fill = tail_hw > 0 and i > 0 and dy.shape[1] <= tail_hw
                if fill:
                    dy_ready = torch.cuda.Event()
                    dy_ready.record(main)
                elif i == 0 and xin.dtype == torch.float32:  # xin = the NCHW image (conv_image path)
                    ops.conv_image_wgrad(xin, dy, L.cout, L.k, L.pad, 1, packs.dwp[id(L.conv)])
                elif i == 0 and self._flat_convs():
                    wgrad(xin, dy, L.cout, 1, 1, 0, packs.dwp[id(L.conv)], on_side=tail_hw == 0)  # xin = im2col patches
                else:
                    wgrad(xin, dy, L.cout, L.k, L.stride, L.pad, packs.dwp[id(L.conv)], on_side=tail_hw == 0)
                if i > 0:
                    prev = grads[i - 1]
                    # this dgrad completes grads[i-1]; when that is the activation gradient of a conv+BN layer the
                    # first pass of its BN backward (sum dz, sum dz*xhat) is folded into the epilogue
                    # (with 32-column epilogue blocks the HBM-bound 1x1 gradients lost more than the separate pass costs:
                    # 91 us fused vs 36 + 45 us for 256->128 @52x52; with 64-column blocks the fused form is 75 us).
                    # The one-launch stride-2 gradients carry it too: +0.39 ms in the three launches for 0.54 ms of
                    # separate reduction passes, 24.72 -> 24.52 ms per step
                    wd2s = self._packs.wpk_d2s.get(id(L.conv)) if prev is None else None
                    use_d2s = wd2s is not None and ops.d2s_dgrad_ok(L.cin, L.k, L.stride, L.pad, dy.shape[2]) and \
                        xin.shape[1] == 2 * dy.shape[1] and xin.shape[2] == 2 * dy.shape[2]
                    owner = self._bn_owner(i - 1) if (fuse and not self.split and (L.stride == 1 or use_d2s) and
                                                      (L.k > 1 or fuse >= 2)) else None
                    if owner is not None and use_d2s:
                        yo = saved[owner][1]  # the depth-to-space form reads y as a contiguous NHWC tensor
                        if not (yo.is_contiguous() and yo.shape[-1] == L.cin and (L.cin & (L.cin - 1)) == 0):
                            owner = None
                    bn_red = None
                    if owner is not None:
                        T = self.layers[owner]
                        bn_red = (saved[owner][1], *saved[owner][2], T.act, T.slope, T.bstats)
                        reduced.add(owner)
                    if use_d2s:
                        dx = ops.conv_dgrad_d2s(dy, wd2s, L.cin, bn_reduce=bn_red)  # one launch, not four parity classes
                    else:
                        dx = ops.conv_dgrad(dy, L.wpk_t, L.cin, L.k, L.stride, L.pad, 1, (xin.shape[1], xin.shape[2]),
                                            out=prev, residual=prev, bn_reduce=bn_red)
                    grads[i - 1] = dx
                    if fill:
                        wgrad(xin, dy, L.cout, L.k, L.stride, L.pad, packs.dwp[id(L.conv)], ev=dy_ready)
            elif L.type == "maxpool":
                (xin,) = saved[i]
                add_grad(i - 1, ops.maxpool_bwd(xin, G, L.pool_stride))
            elif L.type == "upsample":
                if grads[i - 1] is None:
                    grads[i - 1] = ops.upsample_bwd(G)
                else:
                    ops.upsample_bwd(G, grads[i - 1], accumulate=True)
            elif L.type == "route":
                if len(L.inputs) == 1:
                    add_grad(L.inputs[0], G)
                else:
                    c0 = 0
                    for j in L.inputs:
                        c = ops.channels(outs[j])
                        if grads[j] is None:
                            grads[j] = ops.slice_grad(G, c0, c)
                        else:
                            ops.slice_grad(G, c0, c, into=grads[j])
                        c0 += c
            elif L.type == "shortcut":
                if L.fused_alias:
                    add_grad(i - 1, G)  # the conv adds G to its `post_from` input itself
                else:
                    add_grad(L.inputs[0], G)
                    add_grad(L.inputs[1], G)
            grads[i] = None
        if side is not None:
            main.wait_stream(side)
        keep.clear()
        lo, hi, clo, chi = cut[0]
        packs.unpack_range(clo, chi)
        yield arena.flat[lo:hi]
        return views

    # ------------------------------------------------------------------ public entry points
    def train_forward(self, x, targets):
        self._check_input(x)
        require_cuda(targets, "Darknet.forward(targets)")
        with torch.cuda.device(x.device):  # every launch below targets the tensors' device, whatever is current
            return self._train_forward(x, targets)

    def _train_forward(self, x, targets):
        targets = targets.float()
        grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.params)
        if grad and self.model.training and _graphs_enabled() and not torch.cuda.is_current_stream_capturing():
            key = (tuple(x.shape), tuple(targets.shape), str(x.device))
            entry = self._graphs.get(key, 0)
            if isinstance(entry, _GraphedStep):
                # a replay re-uses the static activations of the previous one: while an earlier replayed forward can
                # still be back-propagated (micro-batches whose losses are summed, probe forwards) this call takes the
                # eager launches, which keep per-call state
                if entry.usable() and not self._backward_pending():
                    return self._track(_DarknetGraphFn.apply(entry, x, targets, *self.params))
            elif entry >= GRAPH_WARMUP_CALLS:
                try:
                    if not self._backward_pending():
                        self._graphs[key] = _GraphedStep(self, x, targets)
                        return self._track(_DarknetGraphFn.apply(self._graphs[key], x, targets, *self.params))
                except Exception as e:  # capture is an optimisation: fall back to the eager launches, loudly
                    import warnings

                    warnings.warn(f"b200cv: CUDA-graph capture failed ({e}); staying on eager launches")
                    self._graphs[key] = -(10 ** 9)
                    torch.cuda.synchronize()
            else:
                self._graphs[key] = entry + 1
        private = grad and self._backward_pending()
        out = _DarknetTrainFn.apply(self, x, targets, self.model.training, torch.is_grad_enabled(), private,
                                    *self.params)
        return out if (private or not grad) else self._track(out)

    # A training forward whose backward has not run yet owns the engine's static state (per-layer BatchNorm vectors,
    # the CUDA-graph activations, the gradient arena).  Further grad-enabled forwards issued meanwhile (micro-batches
    # whose losses are summed, probe forwards) run eagerly with private state.
    def _track(self, out):
        self._pending = weakref.ref(out)
        return out

    def _backward_pending(self) -> bool:
        ref = getattr(self, "_pending", None)
        return ref is not None and ref() is not None and ref().grad_fn is not None

    @torch.no_grad()
    def detect(self, x):
        self._check_input(x)
        with torch.cuda.device(x.device):
            det, _ = self._run_forward(x.float(), None, bn_train=self.model.training, want_grad=False)
        return det


class _DarknetTrainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine, x, targets, bn_train, grad_enabled, private, *params):
        # eval-mode BatchNorm (running statistics folded into the conv epilogues) has no backward here: such a pass
        # is computed like the reference computes it and returned as a constant (autograd refuses to differentiate it)
        want_grad = grad_enabled and bn_train and any(ctx.needs_input_grad[6:])
        out7, state = engine._run_forward(x.float(), targets, bn_train=bn_train, want_grad=want_grad, private=private)
        ctx.engine = engine
        ctx.state = state if want_grad else None
        ctx.dev = x.device
        if not want_grad:
            ctx.mark_non_differentiable(out7)
        return out7

    @staticmethod
    def backward(ctx, g7):
        with torch.cuda.device(ctx.dev):
            views = ctx.engine._run_backward(ctx.state, g7)
        if not ctx.state[2]:
            ctx.engine._pending = None
        ctx.state = None
        return (None, None, None, None, None, None, *views)


def _count_launches(n: int):
    from .lib import lib

    lib().launches += n


class _GraphedStep:
    """One training step (fixed shapes) captured as two CUDA graphs -- forward+loss and backward -- sharing a
    memory pool, so the ~1200 launches of a Darknet-53 step cost two graph launches on the host.  Inputs are
    copied into static buffers; the gradient all-reduce stays outside the graph."""

    def __init__(self, engine, x, targets):
        self.engine = engine
        dev = x.device
        self.static_x = x.detach().float().clone()
        self.static_t = targets.detach().clone()
        self.static_g = torch.zeros(7, dtype=torch.float32, device=dev)
        self.static_g[0] = 1.0
        engine._setup(dev)
        self.weight_ptrs = tuple(p.data_ptr() for p in engine.params)
        torch.cuda.synchronize()
        self.pool = torch.cuda.graph_pool_handle()
        self.fwd_graph = torch.cuda.CUDAGraph()
        from .lib import lib

        n0 = lib().launches
        with torch.no_grad():
            with torch.cuda.graph(self.fwd_graph, pool=self.pool):
                self.out7, self.state = engine._run_forward(self.static_x, self.static_t, bn_train=True, want_grad=True)
            n1 = lib().launches
            # backward: one graph per gradient bucket (a single one without data parallelism), so that the all-reduce
            # of a finished bucket can be started between two replays and overlap the rest of the pass
            self.bwd_graphs, self.buckets = [], []
            nbuckets = parallel.grad_buckets()
            gen = engine._run_backward_segments(self.state, self.static_g, True, nbuckets)
            for _ in range(len(engine._bucket_plan(nbuckets))):  # one yield per planned bucket
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=self.pool):
                    self.buckets.append(next(gen))
                self.bwd_graphs.append(g)
            try:  # nothing is launched after the last yield: finish the generator outside any capture
                next(gen)
                raise RuntimeError("backward pass yielded more buckets than planned")
            except StopIteration as fin:
                self.views = fin.value
        # kernel-launching ABI calls recorded in each graph: a replay launches that many of our kernels
        self.fwd_launches, self.bwd_launches = n1 - n0, lib().launches - n1
        self.generation = 0        # bumped by every replayed forward (a stale backward is refused)
        torch.cuda.synchronize()

    def usable(self) -> bool:
        e = self.engine
        return (tuple(p.data_ptr() for p in e.params) == self.weight_ptrs and e._arena is not None
                and not e._arena.aliased_by_param_grads())


class _DarknetGraphFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, step, x, targets, *params):
        step.static_x.copy_(x, non_blocking=True)
        step.static_t.copy_(targets, non_blocking=True)
        step.fwd_graph.replay()
        _count_launches(step.fwd_launches)
        for L in step.engine.layers:  # the replay moved the running statistics: drop the folded inference affines
            L.eval_key = None
        ctx.step = step
        step.generation += 1
        ctx.generation = step.generation
        return step.out7.clone()

    @staticmethod
    def backward(ctx, g7):
        step = ctx.step
        if ctx.generation != step.generation:
            raise RuntimeError("b200cv: this training forward was replayed from a CUDA graph and a LATER forward of the "
                               "same shapes has overwritten its saved activations; back-propagate each forward before "
                               "the next one, or set B200CV_CUDA_GRAPH=0")
        step.engine._pending = None
        with torch.cuda.device(step.static_g.device):
            step.static_g.copy_(g7)
            handles = []
            for graph, chunk in zip(step.bwd_graphs, step.buckets):
                graph.replay()
                handles.append(parallel.allreduce_gradients_async(chunk))  # overlaps the next segment's replay
            _count_launches(step.bwd_launches)
            parallel.finish_allreduces(handles)
        return (None, None, None, *step.views)
