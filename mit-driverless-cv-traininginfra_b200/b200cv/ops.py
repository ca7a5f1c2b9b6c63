"""Tensor-level wrappers over the C ABI (one function per entry point of include/b200cv.h).

PyTorch supplies device memory (torch.empty), the current stream and nothing else: every function
here passes raw device pointers / sizes / strides to libb200cv.so.  Activations are NHWC bf16
tensors of shape [N,H,W,C] (C padded by pad_channels), per-channel vectors are fp32.
"""
from __future__ import annotations

import contextlib
import ctypes
import os
import threading

import torch

from .lib import ACT_LEAKY, ACT_NONE, ACT_RELU, DT_BF16, DT_F32, ConvArgs, lib, ptr, require_cuda, stream_ptr

__all__ = [
    "ACT_NONE", "ACT_LEAKY", "ACT_RELU", "pad_channels", "conv_out_size", "nchw_to_nhwc", "nhwc_to_nchw",
    "pack_weights", "unpack_wgrad", "conv_fwd", "conv_dgrad", "conv_wgrad", "bn_finalize", "bn_apply_act",
    "bn_bwd_reduce", "bn_bwd_finalize", "bn_bwd_apply", "bn_stats_apply_act", "bn_bwd_stats_apply", "act_bwd", "copy_slice", "col_sum", "maxpool_fwd",
    "maxpool_bwd", "upsample_fwd", "upsample_bwd", "act_code", "sm_count", "stats_buffer", "stats_value", "im2col_nchw",
    "use_flat_path", "d2s_dgrad_ok", "conv_dgrad_d2s", "use_image_path", "conv_image_fwd", "conv_image_wgrad", "conv_image_wgrad_bn", "flat_k", "split_mode", "precision", "set_default_precision", "default_split", "channels",
    "bias_grad", "concat_channels", "slice_grad", "split_from_f32", "split_to_f32",
]

# ------------------------------------------------------------------------------ precision mode
# "bf16" (default): NHWC bf16 activations / bf16 tensor-core operands, fp32 accumulation.
# "fp32" (the fp32-parity mode, B200CV_PRECISION=fp32): every activation and weight is stored SPLIT as three bf16
# pieces (p0 = bf16(v), p1 = bf16(v - p0), p2 = bf16(v - p0 - p1): 24 mantissa bits; a tensor is [N,H,W,3C] =
# [p0(C) | p1(C) | p2(C)]), convolutions run six tensor-core passes over the piece products, element-wise kernels
# compute on the summed pieces in fp32 (include/b200cv.h).
# The mode is a thread-local switch set by the engines around their launches (autograd runs backward on its own thread).
_tls = threading.local()
_default_split = os.environ.get("B200CV_PRECISION", "bf16").lower() in ("fp32", "f32", "split", "bf16x3")


def set_default_precision(name: str) -> None:
    """Precision of engines created from now on: "bf16" or "fp32" (the bf16x3 split mode)."""
    global _default_split
    if name not in ("bf16", "fp32"):
        raise ValueError("precision must be 'bf16' or 'fp32'")
    _default_split = name == "fp32"


def default_split() -> bool:
    return _default_split


def split_mode() -> bool:
    return getattr(_tls, "split", False)


@contextlib.contextmanager
def precision(split: bool):
    old = split_mode()
    _tls.split = bool(split)
    try:
        yield
    finally:
        _tls.split = old


def split_pieces() -> int:
    """bf16 pieces per value of the split layout (3)."""
    return int(lib().cdll.b200cv_split_pieces())


def channels(t: torch.Tensor) -> int:
    """Logical channel count of an activation tensor ([N,H,W,C], or [N,H,W,3C] in the split mode)."""
    return t.shape[-1] // split_pieces() if split_mode() else t.shape[-1]


def _mul() -> int:
    return split_pieces() if split_mode() else 1


def pad_channels(c: int) -> int:
    return lib().pad_channels(c)


def act_code(name: str) -> int:
    return {"leaky": ACT_LEAKY, "ReLU": ACT_RELU, "relu": ACT_RELU, "linear": ACT_NONE, "none": ACT_NONE}[name]


def conv_out_size(h: int, k: int, stride: int, pad: int, dil: int = 1) -> int:
    return (h + 2 * pad - dil * (k - 1) - 1) // stride + 1


def _rows(t: torch.Tensor) -> int:
    return t.numel() // t.shape[-1]


# ------------------------------------------------------------------------------ layout / packing
def nchw_to_nhwc(x: torch.Tensor, cpad: int | None = None) -> torch.Tensor:
    """fp32 NCHW -> bf16 NHWC with zero-padded channels."""
    require_cuda(x, "nchw_to_nhwc")
    x = x.contiguous().float()
    n, c, h, w = x.shape
    cpad = cpad or pad_channels(c)
    out = torch.empty(n, h, w, _mul() * cpad, dtype=torch.bfloat16, device=x.device)
    name = "b200cv_nchw_f32_to_nhwc_split" if split_mode() else "b200cv_nchw_f32_to_nhwc_bf16"
    lib().call(name, ptr(x), ptr(out), n, c, h, w, cpad, stream_ptr())
    return out


def nhwc_to_nchw(x: torch.Tensor, c: int) -> torch.Tensor:
    n, h, w, _ = x.shape
    cpad = channels(x)
    out = torch.empty(n, c, h, w, dtype=torch.float32, device=x.device)
    name = "b200cv_nhwc_split_to_nchw_f32" if split_mode() else "b200cv_nhwc_bf16_to_nchw_f32"
    lib().call(name, ptr(x), ptr(out), n, c, h, w, cpad, stream_ptr())
    return out


def split_from_f32(src: torch.Tensor) -> torch.Tensor:
    """fp32 [..., C] (C % 8 == 0) -> split bf16 [..., 3C]."""
    c = src.shape[-1]
    out = torch.empty(*src.shape[:-1], split_pieces() * c, dtype=torch.bfloat16, device=src.device)
    lib().call("b200cv_split_from_f32", ptr(src), src.stride(-2), ptr(out), out.shape[-1], c, _rows(src), c,
               stream_ptr())
    return out


def split_to_f32(src: torch.Tensor) -> torch.Tensor:
    """split bf16 [..., 3C] -> fp32 [..., C] (tests / diagnostics)."""
    c = src.shape[-1] // split_pieces()
    out = torch.empty(*src.shape[:-1], c, dtype=torch.float32, device=src.device)
    lib().call("b200cv_split_to_f32", ptr(src), src.stride(-2), c, ptr(out), c, _rows(src), c, stream_ptr())
    return out


def flat_k(cin: int, k: int) -> int:
    """Padded K of the flat (explicit-im2col) layout of a k x k conv over cin channels."""
    return pad_channels(cin * k * k)


def use_flat_path(cin: int, k: int) -> bool:
    """Few-channel image layers run as explicit im2col + 1x1 conv (one TMA tile per output tile instead of
    k*k nearly-empty 16-channel tiles)."""
    return cin <= 4 and cin * k * k <= 256


def im2col_nchw(x: torch.Tensor, k: int, stride: int, pad: int, dil: int = 1) -> torch.Tensor:
    """NCHW fp32 image -> bf16 patches [N,OH,OW,Kp] (k index = (r*k+s)*C + c)."""
    require_cuda(x, "im2col_nchw")
    x = x.contiguous().float()
    n, c, h, w = x.shape
    oh, ow = conv_out_size(h, k, stride, pad, dil), conv_out_size(w, k, stride, pad, dil)
    kp = flat_k(c, k)
    out = torch.empty(n, oh, ow, _mul() * kp, dtype=torch.bfloat16, device=x.device)
    name = "b200cv_im2col_nchw_f32_split" if split_mode() else "b200cv_im2col_nchw_f32"
    lib().call(name, ptr(x), ptr(out), n, c, h, w, k, k, stride, pad, dil, kp, stream_ptr())
    return out


def use_image_path(cin: int, k: int, stride: int, pad: int, dil: int, cout: int) -> bool:
    """The 3-channel image layers of the bf16 mode run WITHOUT a patch matrix (csrc/conv_image.cu); the fp32-parity
    mode and other shapes keep im2col + 1x1 conv.  B200CV_IMAGE_CONV=0 restores the im2col path."""
    if split_mode() or __import__("os").environ.get("B200CV_IMAGE_CONV", "1") == "0":
        return False
    return bool(lib().cdll.b200cv_conv_image_supported(cin, k, k, stride, pad, dil, cout))


def conv_image_fwd(x: torch.Tensor, w_flat: torch.Tensor, cout: int, k: int, pad: int, dil: int = 1, scale=None,
                   shift=None, act=ACT_NONE, slope=0.0, stats=None) -> torch.Tensor:
    """NCHW fp32 image, flat-packed weights [Cout][1][Kp] -> NHWC bf16 [N,H,W,pad(Cout)] (stride 1, same padding)."""
    require_cuda(x, "conv_image_fwd")
    x = x.contiguous().float()
    n, c, h, w = x.shape
    out = torch.empty(n, h, w, pad_channels(cout), dtype=torch.bfloat16, device=x.device)
    lib().call("b200cv_conv_image_fwd", ptr(x), ptr(w_flat), n, c, h, w, k, k, pad, dil, cout, w_flat.shape[-1],
               ptr(out), out.stride(-2), ptr(scale) if scale is not None else None,
               ptr(shift) if shift is not None else None, act, float(slope),
               ptr(stats) if stats is not None else None, stats.shape[0] if stats is not None else 0, stream_ptr(),
               tag=(c, cout, k, 1, n, h, w))
    return out


def conv_image_wgrad(x: torch.Tensor, dy: torch.Tensor, cout: int, k: int, pad: int, dil: int, out: torch.Tensor):
    """Accumulates the flat packed fp32 gradient [Cout][1][Kp] of an image layer; `out` must be zeroed."""
    x = x.contiguous().float()
    n, c, h, w = x.shape
    lib().call("b200cv_conv_image_wgrad", ptr(x), ptr(dy), dy.stride(-2), n, c, h, w, k, k, pad, dil, cout,
               out.shape[-1], ptr(out), stream_ptr(), tag=(c, cout, k, 1, n, h, w))
    return out


def conv_image_wgrad_bn(x, da, y, partials, count, gamma, coef, dgamma, dbeta, scale, shift, mean, rstd, act, slope,
                        cout: int, k: int, pad: int, dil: int, out: torch.Tensor):
    """conv_image_wgrad with the layer's own BatchNorm + activation backward applied on the fly (da = dL/da): replaces
    bn_bwd_stats_apply + conv_image_wgrad; dy is never written.  Also writes coef / d(gamma) / d(beta)."""
    x = x.contiguous().float()
    n, c, h, w = x.shape
    lib().call("b200cv_conv_image_wgrad_bn", ptr(x), ptr(da), da.stride(-2), ptr(y), y.stride(-2), ptr(partials),
               partials.shape[0], int(count), ptr(gamma), ptr(scale), ptr(shift), ptr(mean), ptr(rstd), act, float(slope),
               ptr(coef), ptr(dgamma), ptr(dbeta), n, c, h, w, k, k, pad, dil, cout, out.shape[-1], ptr(out), stream_ptr(),
               tag=(c, cout, k, 1, n, h, w))
    return out


def pack_weights(w: torch.Tensor, transpose: bool) -> torch.Tensor:
    """OIHW fp32 -> bf16 [O][R*S][Ipad] (forward) or [I][R*S][Opad] (data gradient)."""
    require_cuda(w, "pack_weights")
    o, i, r, s = w.shape
    ipad, opad = pad_channels(i), pad_channels(o)
    m = _mul()
    shape = (i, r * s, m * opad) if transpose else (o, r * s, m * ipad)
    out = torch.empty(shape, dtype=torch.bfloat16, device=w.device)
    lib().call("b200cv_pack_weights", ptr(w.detach().contiguous()), ptr(out), o, i, r, s, ipad, opad,
               int(transpose) | (8 if split_mode() else 0), stream_ptr())
    return out


def unpack_wgrad(dw_packed: torch.Tensor, out_oihw: torch.Tensor):
    o, i, r, s = out_oihw.shape
    lib().call("b200cv_unpack_wgrad", ptr(dw_packed), ptr(out_oihw), o, i, r, s, dw_packed.shape[-1], stream_ptr())


# ------------------------------------------------------------------------------ convolution
def _conv_args(x, wpk, y, cout, k, stride, pad, dil, y_strides, y_dtype, scale, shift, residual, r_strides, act,
               slope, res_after_act, stats) -> ConvArgs:
    a = ConvArgs()
    a.N, a.H, a.W, _ = x.shape
    a.Cin = channels(x)
    if split_mode():
        a.x_lo = a.Cin
        a.y_lo = y.shape[-1] // split_pieces() if y.dtype == torch.bfloat16 else 0
        a.r_lo = residual.shape[-1] // split_pieces() if residual is not None else 0
    a.Cout = cout
    a.R = a.S = k
    a.stride, a.pad, a.dil = stride, pad, dil
    a.x, a.w, a.y = ptr(x), ptr(wpk), ptr(y)
    a.y_dtype = y_dtype
    a.y_sn, a.y_sh, a.y_sw, a.y_sc = y_strides
    a.scale, a.shift, a.residual = ptr(scale), ptr(shift), ptr(residual)
    if residual is not None:
        a.r_sn, a.r_sh, a.r_sw, a.r_sc = r_strides
    a.act, a.slope, a.res_after_act = act, slope, int(res_after_act)
    a.stats = ptr(stats)
    a.stats_parts = stats.shape[0] if stats is not None and stats.dim() == 3 else 1
    return a


def conv_fwd(x, wpk, cout, k, stride, pad, dil=1, out=None, out_dtype=torch.bfloat16, out_channels=None, scale=None,
             shift=None, residual=None, act=ACT_NONE, slope=0.0, res_after_act=False, stats=None,
             nchw_out=False) -> torch.Tensor:
    """x NHWC bf16 -> y.  `out` may be a channel slice (a view whose last-dim stride is 1)."""
    n, h, w, _ = x.shape
    oh, ow = conv_out_size(h, k, stride, pad, dil), conv_out_size(w, k, stride, pad, dil)
    if out is None:
        if nchw_out:
            out = torch.empty(n, cout, oh, ow, dtype=out_dtype, device=x.device)
        else:
            m = _mul() if out_dtype == torch.bfloat16 else 1
            out = torch.empty(n, oh, ow, m * (out_channels or pad_channels(cout)), dtype=out_dtype, device=x.device)
    if nchw_out:
        ys = (out.stride(0), out.stride(2), out.stride(3), out.stride(1))
    else:
        ys = (out.stride(0), out.stride(1), out.stride(2), out.stride(3))
    rs = None
    if residual is not None:
        rs = (residual.stride(0), residual.stride(1), residual.stride(2), residual.stride(3))
    a = _conv_args(x, wpk, out, cout, k, stride, pad, dil, ys, DT_F32 if out.dtype == torch.float32 else DT_BF16,
                   scale, shift, residual, rs, act, slope, res_after_act, stats)
    lib().call("b200cv_conv_fwd", ctypes.byref(a), stream_ptr(), tag=(channels(x), cout, k, stride, n, oh, ow))
    return out


def _set_bn_reduce(a: ConvArgs, bn_reduce):
    by, bsc, bsh, bmean, brstd, bact, bslope, bsums = bn_reduce
    a.bn_sums, a.bn_parts = ptr(bsums), bsums.shape[0]
    a.bn_y, a.bn_y_ld = ptr(by), by.stride(-2)
    a.bn_scale, a.bn_shift, a.bn_mean, a.bn_rstd = ptr(bsc), ptr(bsh), ptr(bmean), ptr(brstd)
    a.bn_act, a.bn_slope = bact, float(bslope)


def conv_dgrad(dy, wpk_t, cin_fwd, k, stride, pad, dil, out_hw, out=None, residual=None, bn_reduce=None) -> torch.Tensor:
    """dy NHWC bf16 [N,OH,OW,pad(Cout)] -> dx NHWC bf16 [N,H,W,pad(Cin_fwd)] (+ residual).
    bn_reduce = (y, scale, shift, mean, rstd, act, slope, sums): also accumulate the BN-backward sums of the layer
    whose activation gradient dx is (stride-1 only), see b200cv_conv_args.bn_sums."""
    n = dy.shape[0]
    h, w = out_hw
    if out is None:
        cpad = pad_channels(cin_fwd)
        out = torch.empty(n, h, w, _mul() * cpad, dtype=torch.bfloat16, device=dy.device)
        if cpad != cin_fwd:
            out.zero_()
    ys = (out.stride(0), out.stride(1), out.stride(2), out.stride(3))
    rs = None
    if residual is not None:
        rs = (residual.stride(0), residual.stride(1), residual.stride(2), residual.stride(3))
    a = _conv_args(dy, wpk_t, out, cin_fwd, k, stride, pad, dil, ys, DT_BF16, None, None, residual, rs, ACT_NONE, 0.0,
                   False, None)
    if bn_reduce is not None:
        _set_bn_reduce(a, bn_reduce)
    lib().call("b200cv_conv_dgrad", ctypes.byref(a), h, w, stream_ptr(),
               tag=(cin_fwd, channels(dy), k, stride, n, dy.shape[1], dy.shape[2]))
    return out


def d2s_dgrad_ok(cin_fwd: int, k: int, stride: int, pad: int, dy_w: int) -> bool:
    """Shapes the one-launch stride-2 data gradient (b200cv_conv_dgrad_d2s) takes; the operand it needs is 16/9 of the
    filter, so it is used for the narrow (HBM-bound) layers only."""
    return (not split_mode() and k == 3 and stride == 2 and pad == 1 and cin_fwd == pad_channels(cin_fwd)
            and cin_fwd % 32 == 0 and cin_fwd <= 128 and dy_w >= 32 and dy_w % 4 == 0)


def conv_dgrad_d2s(dy, wpk_d2s, cin_fwd, out=None, bn_reduce=None) -> torch.Tensor:
    """3x3 stride-2 pad-1 data gradient in one launch: dy NHWC bf16 [N,OH,OW,pad(Cout)] -> dx [N,2*OH,2*OW,Cin_fwd].
    bn_reduce: as for conv_dgrad (the BN input y must be a contiguous NHWC tensor)."""
    n, oh, ow, _ = dy.shape
    h, w = 2 * oh, 2 * ow
    if out is None:
        out = torch.empty(n, h, w, cin_fwd, dtype=torch.bfloat16, device=dy.device)
    ys = (out.stride(0), out.stride(1), out.stride(2), out.stride(3))
    a = _conv_args(dy, wpk_d2s, out, cin_fwd, 3, 2, 1, 1, ys, DT_BF16, None, None, None, None, ACT_NONE, 0.0, False, None)
    if bn_reduce is not None:
        _set_bn_reduce(a, bn_reduce)
    lib().call("b200cv_conv_dgrad_d2s", ctypes.byref(a), h, w, stream_ptr(),
               tag=(cin_fwd, channels(dy), 3, 2, n, oh, ow))
    return out


def conv_wgrad(x, dy, cout, k, stride, pad, dil=1, out=None) -> torch.Tensor:
    """Accumulates into (and returns) the packed fp32 gradient [Cout][k*k][Cin_pad]; `out` must be zeroed."""
    n, h, w, _ = x.shape
    cin = channels(x)
    dwp = out if out is not None else torch.zeros(cout, k * k, cin, dtype=torch.float32, device=x.device)
    sp = split_mode()
    lib().call("b200cv_conv_wgrad", ptr(x), ptr(dy), ptr(dwp), n, h, w, cin, cout, dy.shape[-1], k, k, stride, pad, dil,
               cin if sp else 0, dy.shape[-1] // split_pieces() if sp else 0, stream_ptr(),
               tag=(cin, cout, k, stride, n, dy.shape[1], dy.shape[2]))
    return dwp


# ------------------------------------------------------------------------------ batch-norm + activation
def sm_count(device=None) -> int:
    return torch.cuda.get_device_properties(device if device is not None else torch.cuda.current_device()).multi_processor_count


# rows of a partial-statistics matrix: CTA / block b adds into row b % STAT_PARTS (spreads the atomics)
STAT_PARTS = int(__import__("os").environ.get("B200CV_STAT_PARTS", "1"))
STAT_WORDS = 2  # int64 words of one entry (b200cv_stat: value = w1 * 2^-20 + w2 * 2^-70)


def stats_buffer(cout: int, device) -> torch.Tensor:
    """Zeroed [STAT_PARTS][2*Cout] matrix of b200cv_stat entries (int64 pairs) for conv_fwd(stats=...) /
    bn_bwd_reduce(partials=...): the kernels ADD into it with integer atomics (order-independent, hence
    bit-reproducible), the finalize kernels sum the rows."""
    return torch.zeros(STAT_PARTS, 2 * cout, STAT_WORDS, dtype=torch.int64, device=device)


def stats_encode(values: torch.Tensor) -> torch.Tensor:
    """[1][n] statistics matrix holding the given totals (tests: feed bn_finalize without a conv)."""
    v = values.detach().double().flatten()
    w1 = torch.trunc(v * 2.0 ** 20)
    w2 = torch.round((v - w1 * 2.0 ** -20) * 2.0 ** 70)
    return torch.stack([w1, w2], -1).to(torch.int64).unsqueeze(0).contiguous()


def stats_value(stats: torch.Tensor) -> torch.Tensor:
    """[2*Cout] fp64 totals of a statistics matrix (tests / diagnostics)."""
    s = stats.reshape(-1, stats.shape[-2], STAT_WORDS).sum(0).double()
    return s[:, 0] * 2.0 ** -20 + s[:, 1] * 2.0 ** -70


def bn_finalize(stats, count, gamma, beta, conv_bias, eps, momentum, running_mean, running_var, scale, shift, mean,
                rstd):
    parts = stats.shape[0] if stats.dim() == 3 else 1
    lib().call("b200cv_bn_finalize", ptr(stats), parts, int(count), ptr(gamma), ptr(beta), ptr(conv_bias), float(eps),
               float(momentum), ptr(running_mean), ptr(running_var), ptr(scale), ptr(shift), ptr(mean), ptr(rstd),
               gamma.numel(), stream_ptr())


def bn_apply_act(y, scale, shift, act, slope, out=None, y2=None, scale2=None, shift2=None, post=None):
    c = channels(y)
    if out is None:
        out = torch.empty_like(y)
    if split_mode():
        lib().call("b200cv_split_bn_apply_act", ptr(y), y.stride(-2), ptr(scale), ptr(shift), ptr(y2),
                   0 if y2 is None else y2.stride(-2), ptr(scale2), ptr(shift2), ptr(post),
                   0 if post is None else post.stride(-2), ptr(out), out.stride(-2), _rows(y), c, c, act, float(slope),
                   stream_ptr(), tag=(_rows(y), c))
        return out
    lib().call("b200cv_bn_apply_act", ptr(y), y.stride(-2), ptr(scale), ptr(shift), ptr(y2),
               0 if y2 is None else y2.stride(-2), ptr(scale2), ptr(shift2), ptr(post),
               0 if post is None else post.stride(-2), ptr(out), out.stride(-2), _rows(y), c, act, float(slope),
               stream_ptr(), tag=(_rows(y), c))
    return out


def bn_stats_apply_act(stats, count, gamma, beta, conv_bias, eps, momentum, running_mean, running_var, scale, shift,
                       mean, rstd, y, act, slope, out=None, post=None):
    """bn_finalize + bn_apply_act in ONE launch (every block folds the partial statistics itself)."""
    if split_mode():  # fp32-parity mode: the two un-fused launches
        bn_finalize(stats, count, gamma, beta, conv_bias, eps, momentum, running_mean, running_var, scale, shift, mean,
                    rstd)
        return bn_apply_act(y, scale, shift, act, slope, out=out, post=post)
    c = y.shape[-1]
    if out is None:
        out = torch.empty_like(y)
    parts = stats.shape[0] if stats.dim() == 3 else 1
    lib().call("b200cv_bn_stats_apply_act", ptr(stats), parts, int(count), ptr(gamma), ptr(beta), ptr(conv_bias),
               float(eps), float(momentum), ptr(running_mean), ptr(running_var), ptr(scale), ptr(shift), ptr(mean),
               ptr(rstd), ptr(y), y.stride(-2), ptr(post), 0 if post is None else post.stride(-2), ptr(out),
               out.stride(-2), _rows(y), c, act, float(slope), stream_ptr(), tag=(_rows(y), c))
    return out


def bn_bwd_stats_apply(partials, count, gamma, coef, dgamma, dbeta, da, y, scale, shift, mean, rstd, act, slope,
                       out=None):
    """bn_bwd_finalize + bn_bwd_apply in ONE launch."""
    if split_mode():  # fp32-parity mode: the two un-fused launches
        bn_bwd_finalize(partials, gamma, rstd, count, coef, dgamma, dbeta)
        return bn_bwd_apply(da, y, None, scale, shift, mean, rstd, coef, act, slope, out=out)
    if out is None:
        out = torch.empty_like(y)
    lib().call("b200cv_bn_bwd_stats_apply", ptr(partials), partials.shape[0], int(count), ptr(gamma), ptr(coef),
               ptr(dgamma), ptr(dbeta), ptr(da), da.stride(-2), ptr(y), y.stride(-2), ptr(scale), ptr(shift),
               ptr(mean), ptr(rstd), ptr(out), out.stride(-2), _rows(y), y.shape[-1], act, float(slope),
               stream_ptr(), tag=(_rows(y), y.shape[-1]))
    return out


def bn_bwd_reduce(da, y, aout, scale, shift, mean, rstd, act, slope, partials=None):
    """Adds the per-block sums into `partials` ([nparts][2C], zeroed by the caller; allocated if None)."""
    rows, c = _rows(y), channels(y)
    if partials is None:
        partials = stats_buffer(c, y.device)
    if split_mode():
        lib().call("b200cv_split_bn_bwd_reduce", ptr(da), da.stride(-2), ptr(y), y.stride(-2), ptr(aout),
                   0 if aout is None else aout.stride(-2), ptr(scale), ptr(shift), ptr(mean), ptr(rstd), ptr(partials),
                   partials.shape[0], rows, c, c, act, float(slope), stream_ptr(), tag=(rows, c))
        return partials
    lib().call("b200cv_bn_bwd_reduce", ptr(da), da.stride(-2), ptr(y), y.stride(-2), ptr(aout),
               0 if aout is None else aout.stride(-2), ptr(scale), ptr(shift), ptr(mean), ptr(rstd), ptr(partials),
               partials.shape[0], rows, c, act, float(slope), stream_ptr(), tag=(rows, c))
    return partials


def bn_bwd_finalize(partials, gamma, rstd, count, coef, dgamma, dbeta):
    lib().call("b200cv_bn_bwd_finalize", ptr(partials), partials.shape[0], ptr(gamma), ptr(rstd), int(count),
               ptr(coef), ptr(dgamma), ptr(dbeta), gamma.numel(), stream_ptr())


def bn_bwd_apply(da, y, aout, scale, shift, mean, rstd, coef, act, slope, out=None):
    if out is None:
        out = torch.empty_like(y)
    if split_mode():
        c = channels(y)
        lib().call("b200cv_split_bn_bwd_apply", ptr(da), da.stride(-2), ptr(y), y.stride(-2), ptr(aout),
                   0 if aout is None else aout.stride(-2), ptr(scale), ptr(shift), ptr(mean), ptr(rstd), ptr(coef),
                   ptr(out), out.stride(-2), _rows(y), c, c, act, float(slope), stream_ptr(), tag=(_rows(y), c))
        return out
    lib().call("b200cv_bn_bwd_apply", ptr(da), da.stride(-2), ptr(y), y.stride(-2), ptr(aout),
               0 if aout is None else aout.stride(-2), ptr(scale), ptr(shift), ptr(mean), ptr(rstd), ptr(coef),
               ptr(out), out.stride(-2), _rows(y), y.shape[-1], act, float(slope), stream_ptr(),
               tag=(_rows(y), y.shape[-1]))
    return out


def bn_bwd_reduce2(da, aout, y_a, y_b, mean_a, rstd_a, mean_b, rstd_b, act, slope):
    """BN-backward sums of two BatchNorms that meet in one activation (out = act(bn_a(y_a) + bn_b(y_b))): one pass
    over da / aout for both.  Returns the two zero-initialised-and-filled partial matrices."""
    rows, c = _rows(y_a), channels(y_a)
    pa, pb = stats_buffer(c, y_a.device), stats_buffer(c, y_a.device)
    if split_mode():  # one pass per layer, the ReLU mask taken from the saved output
        bn_bwd_reduce(da, y_a, aout, None, None, mean_a, rstd_a, act, slope, partials=pa)
        bn_bwd_reduce(da, y_b, aout, None, None, mean_b, rstd_b, act, slope, partials=pb)
        return pa, pb
    lib().call("b200cv_bn_bwd_reduce2", ptr(da), da.stride(-2), ptr(aout), aout.stride(-2), ptr(y_a), y_a.stride(-2),
               ptr(y_b), y_b.stride(-2), ptr(mean_a), ptr(rstd_a), ptr(mean_b), ptr(rstd_b), ptr(pa), ptr(pb),
               pa.shape[0], rows, c, act, float(slope), stream_ptr(), tag=(rows, c))
    return pa, pb


def bn_bwd_apply2(da, aout, y_a, y_b, mean_a, rstd_a, mean_b, rstd_b, coef_a, coef_b, act, slope):
    """dy of both layers of bn_bwd_reduce2 in one pass.  Returns (dy_a, dy_b)."""
    if split_mode():
        return (bn_bwd_apply(da, y_a, aout, None, None, mean_a, rstd_a, coef_a, act, slope),
                bn_bwd_apply(da, y_b, aout, None, None, mean_b, rstd_b, coef_b, act, slope))
    dy_a, dy_b = torch.empty_like(y_a), torch.empty_like(y_b)
    lib().call("b200cv_bn_bwd_apply2", ptr(da), da.stride(-2), ptr(aout), aout.stride(-2), ptr(y_a), y_a.stride(-2),
               ptr(y_b), y_b.stride(-2), ptr(mean_a), ptr(rstd_a), ptr(mean_b), ptr(rstd_b), ptr(coef_a), ptr(coef_b),
               ptr(dy_a), dy_a.stride(-2), ptr(dy_b), dy_b.stride(-2), _rows(y_a), y_a.shape[-1], act, float(slope),
               stream_ptr(), tag=(_rows(y_a), y_a.shape[-1]))
    return dy_a, dy_b


def act_bwd(da, aout, act, slope):
    out = torch.empty_like(aout)
    if split_mode():
        c = channels(aout)
        lib().call("b200cv_split_act_bwd", ptr(da), da.stride(-2), ptr(aout), aout.stride(-2), ptr(out), out.stride(-2),
                   _rows(aout), c, c, act, float(slope), stream_ptr())
        return out
    lib().call("b200cv_act_bwd", ptr(da), da.stride(-2), ptr(aout), aout.stride(-2), ptr(out), out.stride(-2),
               _rows(aout), aout.shape[-1], act, float(slope), stream_ptr())
    return out


def copy_slice(src, dst, accumulate=False, c=None, src_lo=None, dst_lo=None):
    """dst[..., :C] (=|+=) src[..., :C] for NHWC views with unit channel stride.  Split mode: whole tensors by
    default (lo = half the width); channel slices of wider buffers pass their hi views plus c / src_lo / dst_lo."""
    if split_mode():
        c = c if c is not None else src.shape[-1] // split_pieces()
        src_lo = src_lo if src_lo is not None else src.shape[-1] // split_pieces()
        dst_lo = dst_lo if dst_lo is not None else dst.shape[-1] // split_pieces()
        lib().call("b200cv_split_copy_slice", ptr(src), src.stride(-2), src_lo, ptr(dst), dst.stride(-2), dst_lo,
                   _rows(src), c, int(accumulate), stream_ptr(), tag=(_rows(src), c))
        return dst
    lib().call("b200cv_copy_slice", ptr(src), src.stride(-2), ptr(dst), dst.stride(-2), _rows(src), src.shape[-1],
               int(accumulate), stream_ptr(), tag=(_rows(src), src.shape[-1]))
    return dst


def concat_channels(parts):
    """torch.cat(parts, dim=channel) of NHWC activations (route layers, CVC-YOLOv3/models.py:322-324)."""
    b, hh, ww = parts[0].shape[:3]
    widths = [channels(p) for p in parts]
    total = sum(widths)
    out = torch.empty(b, hh, ww, _mul() * total, dtype=parts[0].dtype, device=parts[0].device)
    c0 = 0
    for p, c in zip(parts, widths):
        copy_slice(p, out[..., c0:c0 + c], c=c, dst_lo=total)
        c0 += c
    return out


def slice_grad(g, c0, c, into=None):
    """Channels [c0, c0+c) of the gradient of a concat buffer, as a tensor of its own (or accumulated into `into`)."""
    total = channels(g)
    sl = g[..., c0:c0 + c]
    if into is None:
        dst = torch.empty(*g.shape[:3], _mul() * c, dtype=g.dtype, device=g.device)
        return copy_slice(sl, dst, c=c, src_lo=total)
    return copy_slice(sl, into, accumulate=True, c=c, src_lo=total)


def col_sum(x, out):
    lib().call("b200cv_col_sum", ptr(x), x.stride(-2), _rows(x), x.shape[-1], ptr(out), stream_ptr())
    return out


def bias_grad(dy, cout):
    """fp32 [cout] column sums of a gradient tensor (conv bias gradient)."""
    tmp = torch.zeros(dy.shape[-1], dtype=torch.float32, device=dy.device)
    col_sum(dy, tmp)  # split mode: the pieces are summed as separate columns ...
    if split_mode():
        w = dy.shape[-1] // split_pieces()
        return tmp.view(split_pieces(), w)[:, :cout].sum(0)  # ... and added here
    return tmp[:cout]


def maxpool_fwd(x, stride):
    n, h, w, c = x.shape
    oh, ow = (h // 2, w // 2) if stride == 2 else (h, w)
    y = torch.empty(n, oh, ow, c, dtype=x.dtype, device=x.device)
    if split_mode():
        lib().call("b200cv_split_maxpool2x2_fwd", ptr(x), ptr(y), n, h, w, c // split_pieces(), stride, stream_ptr())
        return y
    lib().call("b200cv_maxpool2x2_fwd", ptr(x), ptr(y), n, h, w, c, stride, stream_ptr())
    return y


def maxpool_bwd(x, dy, stride):
    n, h, w, c = x.shape
    dx = torch.empty_like(x)
    if split_mode():
        lib().call("b200cv_split_maxpool2x2_bwd", ptr(x), ptr(dy), ptr(dx), n, h, w, c // split_pieces(), stride,
                   stream_ptr())
        return dx
    lib().call("b200cv_maxpool2x2_bwd", ptr(x), ptr(dy), ptr(dx), n, h, w, c, stride, stream_ptr())
    return dx


def upsample_fwd(x, out=None):
    n, h, w, c = x.shape
    if out is None:
        out = torch.empty(n, 2 * h, 2 * w, c, dtype=x.dtype, device=x.device)
    lib().call("b200cv_upsample2x_fwd", ptr(x), ptr(out), out.stride(-2), n, h, w, c, stream_ptr())
    return out


def upsample_bwd(dy, dx=None, accumulate=False):
    n, oh, ow, c = dy.shape
    if dx is None:
        dx = torch.empty(n, oh // 2, ow // 2, c, dtype=dy.dtype, device=dy.device)
    if split_mode():
        lib().call("b200cv_split_upsample2x_bwd", ptr(dy), ptr(dx), n, oh // 2, ow // 2, c // split_pieces(), int(accumulate),
                   stream_ptr())
        return dx
    lib().call("b200cv_upsample2x_bwd", ptr(dy), dy.stride(-2), ptr(dx), n, oh // 2, ow // 2, c, int(accumulate),
               stream_ptr())
    return dx
