"""The C-ABI library loads without a GPU and exports every symbol include/b200cv.h declares."""
import ctypes
import os
import re
import subprocess

from b200cv import lib as L


def test_header_symbols_exported():
    protos = L.parse_header()
    assert len(protos) >= 30
    cdll = ctypes.CDLL(L.LIB_PATH)
    for name in protos:
        assert hasattr(cdll, name), f"{name} declared in include/b200cv.h but not exported"


def test_exports_match_header():
    out = subprocess.run(["nm", "-D", "--defined-only", L.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (b200cv_\w+)", out))
    assert exported == set(L.parse_header()), exported ^ set(L.parse_header())


def test_version_and_pad_rule():
    lib = L.lib()
    assert "sm_100a" in lib.version()
    assert [lib.pad_channels(c) for c in (1, 3, 16, 17, 18, 32, 33, 64, 255, 256, 1024)] == \
           [16, 16, 16, 32, 32, 32, 64, 64, 256, 256, 1024]


def test_conv_args_struct_matches_header():
    text = open(L.HEADER).read()
    body = text[text.index("typedef struct b200cv_conv_args {"):text.index("} b200cv_conv_args;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split("{", 1)[1].split(";"):
        decl = decl.strip()
        if not decl:
            continue
        names = decl.replace("*", " ").split()
        # "int32_t N, H, W, Cin" -> N H W Cin ; "const void* x" -> x
        first = [n for n in names if n not in ("const", "void", "float", "int32_t", "int64_t")]
        fields += [n.strip(",") for n in first]
    assert fields == [f[0] for f in L.ConvArgs._fields_]


def test_bad_arguments_fail_loudly():
    lib = L.lib()
    rc = lib.cdll.b200cv_conv_fwd(None, None)
    assert rc == -1 and "null" in lib.last_error()


def test_product_has_no_cpu_path():
    import pytest
    import torch

    from b200cv import ops

    with pytest.raises(L.B200CVError):
        ops.nchw_to_nhwc(torch.zeros(1, 3, 4, 4))


def test_new_rows_have_no_cpu_path_either():
    """The 8f rows (NMS / crop / AP kernels, fused optimizers, the pipeline) refuse CPU tensors instead of silently
    computing on the host."""
    import pytest
    import torch

    from b200cv import detect_ops
    from b200cv import optim as boptim

    with pytest.raises(L.B200CVError):
        detect_ops.detect_nms(torch.zeros(1, 10, 6), 0.5, 0.3)
    from utils.nms import nms

    with pytest.raises(L.B200CVError):
        nms(torch.zeros(3, 4), torch.ones(3))
    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.ones(4)
    for opt in (boptim.FusedAdam([p], lr=0.1), boptim.FusedSGD([p], lr=0.1, momentum=0.9)):
        with pytest.raises(L.B200CVError):
            opt.step()
    assert torch.equal(p.detach(), torch.zeros(4))  # nothing was updated on the host


def test_fused_optimizer_host_logic():
    """Constructor validation and torch.optim plumbing that need no GPU: param_groups, schedulers, state_dict keys."""
    import pytest
    import torch

    from b200cv import optim as boptim

    with pytest.raises(ValueError):
        boptim.FusedAdam([torch.nn.Parameter(torch.zeros(1))], lr=-1.0)
    with pytest.raises(ValueError):
        boptim.FusedSGD([torch.nn.Parameter(torch.zeros(1))], momentum=-0.1)
    ps = [torch.nn.Parameter(torch.zeros(3)), torch.nn.Parameter(torch.zeros(2, 2))]
    opt = boptim.FusedAdam(ps, lr=0.1, weight_decay=5e-4)
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=1, gamma=0.95)   # CVC-YOLOv3/train.py:199
    opt.step()          # no gradients yet: a no-op, like torch.optim
    sched.step()
    assert opt.param_groups[0]["lr"] == pytest.approx(0.095)
    assert opt.state_dict()["param_groups"][0]["weight_decay"] == 5e-4
