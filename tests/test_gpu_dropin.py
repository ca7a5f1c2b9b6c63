"""The drop-in claim, executed: the reference's OWN training scripts (unmodified files staged under baseline/_ref by
__graft_entry__.build()) run on top of the B200 modules -- CVC-YOLOv3/train.py `main()` end to end (load_weights with the
255 -> 18 filter truncation, DataLoader, Adam + StepLR, run_epoch, save_weights, the validation-loss pass, validate.py's
mAP pass with the reference's NMS) and RektNet/train_eval.py `train_model` / `eval_model`.  Each script runs in its own
interpreter (both projects have top-level modules called `utils`, and train.py sets process-wide state)."""
import json
import os
import re
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


def _run(driver, work, *extra, cpu=False):
    if not os.path.isfile(os.path.join(REF, "CVC-YOLOv3", "train.py")):
        pytest.skip("baseline/_ref not staged (run __graft_entry__.build() where the reference tree is mounted)")
    os.makedirs(str(work), exist_ok=True)
    env = dict(os.environ)
    env.pop("B200CV_PRECISION", None)
    if cpu:
        env["CUDA_VISIBLE_DEVICES"] = ""
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dropin", driver), ROOT, str(work), *extra],
                       cwd=str(work), env=env, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-3000:] + "\n" + p.stderr[-3000:]
    return json.loads(p.stdout.strip().splitlines()[-1]), p.stdout


def _val_losses(log):
    return [float(m.group(1)) for m in re.finditer(r"Average Validation Loss:\s+([0-9.eE+-]+)", log)]


def test_reference_train_py_runs_unchanged_on_the_b200_modules(tmp_path):
    res, log = _run("yolo_driver.py", tmp_path / "b200")
    assert res["files"] == ["1.weights", "2.weights"]  # model.save_weights at every checkpoint (train.py:217)
    assert 0 < res["val_loss"] < 1e4 and abs(res["val_loss"] - res["result_txt"]) < 1e-4 * res["val_loss"]
    assert "Model in train mode" in log and "Calculating loss on validate data" in log and "mAP:" in log
    assert len(res["epoch_losses"]) == 7 and all(v == v and v > 0 for v in res["epoch_losses"])
    assert res["gpu_launches"] > 500  # the work was done by the B200 kernels, not by a fallback
    # the same script on the reference's OWN models.py (CPU fp32) from the same start weights: same trajectory
    ref_res, ref_log = _run("yolo_driver.py", tmp_path / "ref", "reference", str(tmp_path / "b200" / "start.weights"),
                            cpu=True)
    mine, theirs = _val_losses(log), _val_losses(ref_log)
    assert len(mine) == len(theirs) == 2
    for a, b in zip(mine, theirs):  # six Adam steps in bf16 vs fp32
        assert abs(a - b) <= 2e-2 * b, (mine, theirs)
    ours_w = open(tmp_path / "b200" / "out" / "2.weights", "rb").read()
    ref_w = open(tmp_path / "ref" / "out" / "2.weights", "rb").read()
    assert len(ours_w) == len(ref_w) and ours_w[:20] == ref_w[:20]  # same file layout, same header


def test_reference_train_eval_py_runs_unchanged_on_the_b200_modules(tmp_path):
    res, log = _run("rektnet_driver.py", tmp_path / "b200")
    pat = r"Training: MSE/Geometric/Total Loss: [^/]+/[^/]+/([0-9.eE+-]+)"
    mine = [float(m.group(1)) for m in re.finditer(pat, log)]
    assert len(mine) == 3 and all(v == v and v > 0 for v in res["val_after"])
    assert abs(res["lr"] - 1e-3 * 0.999 ** 3) < 1e-9     # ExponentialLR stepped once per epoch (train_eval.py:85)
    assert "Starting validation" in log
    assert res["gpu_launches"] > 500
    # the same loop on the reference's own modules (CPU fp32): the per-epoch training losses agree
    _, ref_log = _run("rektnet_driver.py", tmp_path / "ref", "reference", cpu=True)
    theirs = [float(m.group(1)) for m in re.finditer(pat, ref_log)]
    assert len(theirs) == 3
    for a, b in zip(mine, theirs):  # twelve Adam steps in bf16 vs fp32
        assert abs(a - b) <= 3e-2 * b, (mine, theirs)
