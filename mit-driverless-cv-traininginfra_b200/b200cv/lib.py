"""ctypes binding of libb200cv.so -- the thin C-ABI layer between the Python model/loss API and the
hand-written sm_100a kernels.

The prototypes are read from ``include/b200cv.h`` (the single source of truth for the ABI), so a
symbol the header declares but the library does not export fails at import time, loudly.  There is
no Python / CPU fallback: if the shared library is missing, importing this module raises.
"""
from __future__ import annotations

import ctypes
import os
import re
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
HEADER = os.path.join(_ROOT, "include", "b200cv.h")
LIB_PATH = os.path.join(_HERE, "libb200cv.so")

ACT_NONE, ACT_LEAKY, ACT_RELU = 0, 1, 2
DT_BF16, DT_F32 = 0, 1


class ConvArgs(ctypes.Structure):
    """Mirror of ``b200cv_conv_args`` (include/b200cv.h)."""

    _fields_ = [
        ("N", ctypes.c_int32), ("H", ctypes.c_int32), ("W", ctypes.c_int32), ("Cin", ctypes.c_int32),
        ("Cout", ctypes.c_int32),
        ("R", ctypes.c_int32), ("S", ctypes.c_int32), ("stride", ctypes.c_int32), ("pad", ctypes.c_int32),
        ("dil", ctypes.c_int32),
        ("x", ctypes.c_void_p), ("w", ctypes.c_void_p), ("y", ctypes.c_void_p),
        ("y_dtype", ctypes.c_int32),
        ("y_sn", ctypes.c_int64), ("y_sh", ctypes.c_int64), ("y_sw", ctypes.c_int64), ("y_sc", ctypes.c_int64),
        ("scale", ctypes.c_void_p), ("shift", ctypes.c_void_p), ("residual", ctypes.c_void_p),
        ("r_sn", ctypes.c_int64), ("r_sh", ctypes.c_int64), ("r_sw", ctypes.c_int64), ("r_sc", ctypes.c_int64),
        ("act", ctypes.c_int32), ("slope", ctypes.c_float), ("res_after_act", ctypes.c_int32),
        ("stats", ctypes.c_void_p), ("stats_parts", ctypes.c_int32),
        ("bn_sums", ctypes.c_void_p), ("bn_parts", ctypes.c_int32), ("bn_y", ctypes.c_void_p),
        ("bn_y_ld", ctypes.c_int64), ("bn_scale", ctypes.c_void_p), ("bn_shift", ctypes.c_void_p),
        ("bn_mean", ctypes.c_void_p), ("bn_rstd", ctypes.c_void_p), ("bn_act", ctypes.c_int32),
        ("bn_slope", ctypes.c_float),
        ("x_lo", ctypes.c_int64), ("y_lo", ctypes.c_int64), ("r_lo", ctypes.c_int64),
    ]


_SCALARS = {
    "int": ctypes.c_int, "int32_t": ctypes.c_int32, "int64_t": ctypes.c_int64, "float": ctypes.c_float,
    "double": ctypes.c_double,
}


def parse_header(path: str = HEADER):
    """Return {name: (restype, [argtypes])} for every ``b200cv_*`` prototype in the header."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    text = re.sub(r"^\s*#.*$", " ", text, flags=re.M)
    protos = {}
    for m in re.finditer(r"(const\s+char\s*\*|int)\s+(b200cv_\w+)\s*\(([^)]*)\)\s*;", text, flags=re.S):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        restype = ctypes.c_char_p if "char" in ret else ctypes.c_int
        argtypes = []
        args = " ".join(args.split())
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                else:
                    ty = a.split()[-2] if len(a.split()) >= 2 else a
                    argtypes.append(_SCALARS[ty])
        protos[name] = (restype, argtypes)
    return protos


class B200CVError(RuntimeError):
    pass


class _Lib:
    def __init__(self):
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C mit-driverless-cv-traininginfra_b200/csrc` (no CPU fallback exists)")
        self.cdll = ctypes.CDLL(LIB_PATH)
        self.protos = parse_header()
        self.launches = 0  # kernel-launching ABI calls made (bench.py reports this)
        self._prof = None
        self._lock = threading.Lock()
        for name, (restype, argtypes) in self.protos.items():
            fn = getattr(self.cdll, name)  # AttributeError => header/library mismatch
            fn.restype = restype
            fn.argtypes = argtypes

    def version(self) -> str:
        return self.cdll.b200cv_version().decode()

    def last_error(self) -> str:
        return self.cdll.b200cv_last_error().decode()

    def call(self, name: str, *args, tag=None):
        prof = self._prof
        if prof is not None:
            import torch

            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        rc = getattr(self.cdll, name)(*args)
        self.launches += 1
        if prof is not None:
            e1.record()
            prof.append((name, e0, e1, tag))
        if rc != 0:
            raise B200CVError(f"{name} failed (rc={rc}): {self.last_error()}")

    def profile_step(self, fn, detail=False):
        """Run fn() with a CUDA-event pair around every ABI call (on the current stream); returns
        {entry point: total device milliseconds}.  Measurement aid for bench.py -- not a hot-path feature."""
        import torch

        torch.cuda.synchronize()
        self._prof = []
        try:
            fn()
            torch.cuda.synchronize()
            if detail:  # every call in launch order: (entry point, tag, ms)
                return [(name, tag, e0.elapsed_time(e1)) for name, e0, e1, tag in self._prof]
            out = {}
            for name, e0, e1, _ in self._prof:
                out[name] = out.get(name, 0.0) + e0.elapsed_time(e1)
        finally:
            self._prof = None
        return out

    def pad_channels(self, c: int) -> int:
        return int(self.cdll.b200cv_pad_channels(int(c)))


_lib = None
_lib_lock = threading.Lock()


def lib() -> _Lib:
    global _lib
    if _lib is None:
        with _lib_lock:
            if _lib is None:
                _lib = _Lib()
    return _lib


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr():
    import torch

    return torch.cuda.current_stream().cuda_stream


def require_cuda(t, what: str):
    if not t.is_cuda:
        raise B200CVError(
            f"{what}: tensor is on {t.device}; the B200 path has no CPU implementation "
            "(oracle/ is test infrastructure only)")
