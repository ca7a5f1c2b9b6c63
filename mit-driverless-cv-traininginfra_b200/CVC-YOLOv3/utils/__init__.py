"""Drop-in overlay of the reference's ``utils`` package.

``models``, ``utils.parse_config`` and the hot-path functions of ``utils.utils`` (build_targets, bbox_iou,
weights_init_normal) are B200-native here.  Everything else the reference scripts import from
``utils`` (datasets, nms, Logger, padding helpers ...) is outside the accelerated path: when the
environment variable B200CV_REFERENCE_ROOT points at a checkout of the reference, those modules are
resolved from ``$B200CV_REFERENCE_ROOT/CVC-YOLOv3/utils`` so train.py / validate.py / detect.py run
unchanged with this directory first on sys.path.
"""
import os as _os

_ref = _os.environ.get("B200CV_REFERENCE_ROOT")
if _ref:
    _p = _os.path.join(_ref, "CVC-YOLOv3", "utils")
    if _os.path.isdir(_p) and _p not in __path__:
        __path__.append(_p)
