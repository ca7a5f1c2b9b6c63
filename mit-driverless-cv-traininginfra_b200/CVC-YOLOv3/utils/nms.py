"""``utils.nms.nms`` of the reference (CVC-YOLOv3/utils/nms.py:4-61) on the B200 kernel.

Same signature and return value: the indices (int64, into ``boxes``) of the kept boxes in visiting order
(descending score).  Equal scores are visited later-index-first -- the order a stable ascending sort gives the
reference's loop; the reference's own (unstable) sort leaves it unspecified.  CUDA tensors only; ``top_k`` <= 512."""
import os as _os
import sys as _sys

import torch

_pkg_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _pkg_root not in _sys.path:
    _sys.path.insert(0, _pkg_root)

from b200cv import detect_ops as _detect_ops  # noqa: E402
from b200cv.lib import require_cuda as _require_cuda  # noqa: E402


def nms(boxes, scores, overlap=0.5, top_k=200):
    _require_cuda(boxes, "nms")
    if boxes.numel() == 0:
        return scores.new_zeros(scores.size(0)).long()  # reference :15-17 returns the zero-filled buffer
    rows = torch.cat([boxes.float(), scores.float().reshape(-1, 1)], 1).contiguous()
    d = _detect_ops.detect_nms(rows, float("-inf"), float(overlap), int(top_k), corners=True)
    n = int(d.counts[0])  # the reference returns a data-dependent length too (keep[:count])
    return d.rows[0, :n].long()
