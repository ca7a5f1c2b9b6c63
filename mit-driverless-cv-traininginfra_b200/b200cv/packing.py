"""Per-network weight staging: every conv weight is re-packed (OIHW fp32 -> bf16 operand layouts) with ONE
multi-tensor launch per step, weight gradients are accumulated in one flat packed fp32 arena (one memset) and
un-packed into the OIHW gradient arena with one more launch."""
from __future__ import annotations


import numpy as np
import torch

from .lib import lib, stream_ptr
from .ops import flat_k, pad_channels

_ENTRY = np.dtype([("src", np.uint64), ("dst", np.uint64), ("O", np.int32), ("I", np.int32), ("RS", np.int32),
                   ("Ipad", np.int32), ("Opad", np.int32), ("transpose", np.int32)])
assert _ENTRY.itemsize == 40


class GradArena:
    """Flat fp32 gradient arena with one view per parameter (what the data-parallel all-reduce sends)."""

    def __init__(self, params, device):
        self.params = params
        sizes = [p.numel() for p in params]
        self.flat = torch.zeros(sum(sizes), dtype=torch.float32, device=device)
        self.views, o = [], 0
        for p, n in zip(params, sizes):
            self.views.append(self.flat[o:o + n].view_as(p))
            o += n
        self.view_of = {id(p): v for p, v in zip(params, self.views)}

    def aliased_by_param_grads(self) -> bool:
        """True if some param.grad still IS a view of this arena (zero_grad(set_to_none=False) style loops):
        writing the arena in place would then corrupt autograd's accumulation."""
        lo = self.flat.data_ptr()
        hi = lo + self.flat.numel() * 4
        return any(p.grad is not None and lo <= p.grad.data_ptr() < hi for p in self.params)


class ConvPackSet:
    def __init__(self, convs, device, grad_arena: GradArena, flat=(), split=False, d2s=()):
        """convs: list of (nn.Conv2d, need_transposed_pack); `flat`: convs that use the explicit-im2col
        layout [O][1][Kp] (k = tap*I + i) instead of [O][taps][Ipad]; `split`: fp32-parity packs, every innermost
        channel run stored as [hi | lo] (the fp32 gradient layout is the same in both modes); `d2s`: 3x3 stride-2
        convs that additionally get the depth-to-space operand of the one-launch data gradient (`wpk_d2s`)."""
        self.convs = [c for c, _ in convs]
        self.device = device
        self._flat = {id(c) for c in flat}
        self.split = bool(split)
        m = int(lib().cdll.b200cv_split_pieces()) if split else 1
        fwd_sizes, t_sizes, g_sizes = [], [], []
        for conv, need_t in convs:
            o, i, r, s = conv.weight.shape
            n_fwd = o * flat_k(i, r) if id(conv) in self._flat else o * r * s * pad_channels(i)
            fwd_sizes.append(m * n_fwd)
            t_sizes.append(m * i * r * s * pad_channels(o) if need_t else 0)
            g_sizes.append(n_fwd)
        al = lambda n: (n + 127) // 128 * 128  # keep every view 256-byte aligned
        self._wpk = torch.empty(sum(al(n) for n in fwd_sizes), dtype=torch.bfloat16, device=device)
        self._wpk_t = torch.empty(max(1, sum(al(n) for n in t_sizes)), dtype=torch.bfloat16, device=device)
        self.dwp_flat = torch.zeros(sum(al(n) for n in g_sizes), dtype=torch.float32, device=device)
        self.wpk, self.wpk_t, self.dwp = {}, {}, {}
        of = ot = og = 0
        for (conv, need_t), nf, nt, ng in zip(convs, fwd_sizes, t_sizes, g_sizes):
            o, i, r, s = conv.weight.shape
            shape = (o, 1, flat_k(i, r)) if id(conv) in self._flat else (o, r * s, pad_channels(i))
            self.wpk[id(conv)] = self._wpk[of:of + nf].view(shape[0], shape[1], m * shape[2])
            of += al(nf)
            if need_t:
                self.wpk_t[id(conv)] = self._wpk_t[ot:ot + nt].view(i, r * s, m * pad_channels(o))
                ot += al(nt)
            else:
                self.wpk_t[id(conv)] = None
            self.dwp[id(conv)] = self.dwp_flat[og:og + ng].view(shape)
            og += al(ng)
        self.wpk_d2s, self._d2s = {}, [c for c in d2s if not split]
        if self._d2s:
            sizes = [16 * c.weight.shape[1] * pad_channels(c.weight.shape[0]) for c in self._d2s]
            self._wpk_d2s = torch.empty(sum(al(n) for n in sizes), dtype=torch.bfloat16, device=device)
            od = 0
            for c, n in zip(self._d2s, sizes):
                self.wpk_d2s[id(c)] = self._wpk_d2s[od:od + n].view(4 * c.weight.shape[1], 4, pad_channels(c.weight.shape[0]))
                od += al(n)
        self._need_t = [t for _, t in convs]
        self._ptrs = None
        self._pack_table = self._pack_table_fwd = None
        self._n_pack = self._n_fwd = 0
        # gradient un-pack table (static: packed arena -> OIHW views of the gradient arena)
        rows = []
        for conv in self.convs:
            o, i, r, s = conv.weight.shape
            fl = id(conv) in self._flat
            rows.append((self.dwp[id(conv)].data_ptr(), grad_arena.view_of[id(conv.weight)].data_ptr(), o, i, r * s,
                         flat_k(i, r) if fl else pad_channels(i), 0, 2 if fl else 0))
        self._unpack_table = self._to_device(rows)
        self._n_unpack = len(rows)

    def _to_device(self, rows):
        arr = np.array(rows, dtype=_ENTRY)
        return torch.from_numpy(arr.view(np.uint8).copy()).to(self.device)

    def _refresh_tables(self):
        ptrs = tuple(c.weight.data_ptr() for c in self.convs)
        if ptrs == self._ptrs:
            return
        self._ptrs = ptrs
        fwd, both = [], []
        for conv, need_t in zip(self.convs, self._need_t):
            o, i, r, s = conv.weight.shape
            fl = id(conv) in self._flat
            sp = 8 if self.split else 0  # b200cv_pack_entry.transpose | 8 = split pack
            e = (conv.weight.data_ptr(), self.wpk[id(conv)].data_ptr(), o, i, r * s,
                 flat_k(i, r) if fl else pad_channels(i), pad_channels(o), (2 if fl else 0) | sp)
            fwd.append(e)
            both.append(e)
            if need_t:
                both.append((conv.weight.data_ptr(), self.wpk_t[id(conv)].data_ptr(), o, i, r * s, pad_channels(i),
                             pad_channels(o), 1 | sp))
        for conv in self._d2s:
            o, i, r, s = conv.weight.shape
            both.append((conv.weight.data_ptr(), self.wpk_d2s[id(conv)].data_ptr(), o, i, r * s, pad_channels(i),
                         pad_channels(o), 3))
        self._pack_table_fwd, self._n_fwd = self._to_device(fwd), len(fwd)
        self._pack_table, self._n_pack = self._to_device(both), len(both)

    def pack_all(self, need_t: bool):
        self._refresh_tables()
        if need_t:
            lib().call("b200cv_pack_weights_multi", self._pack_table.data_ptr(), self._n_pack, stream_ptr())
        else:
            lib().call("b200cv_pack_weights_multi", self._pack_table_fwd.data_ptr(), self._n_fwd, stream_ptr())

    def zero_grads(self):
        self.dwp_flat.zero_()

    def unpack_all(self):
        lib().call("b200cv_unpack_wgrad_multi", self._unpack_table.data_ptr(), self._n_unpack, stream_ptr())

    def unpack_range(self, lo: int, hi: int):
        """Un-pack the weight gradients of convs [lo, hi) (network order) only: lets the data-parallel all-reduce of a
        finished bucket of layers start while backward is still working on the earlier layers."""
        if hi > lo:
            lib().call("b200cv_unpack_wgrad_multi", self._unpack_table.data_ptr() + lo * _ENTRY.itemsize, hi - lo,
                       stream_ptr())
