"""Tile-and-scale input pipeline: the oracle restatement is pinned to the installed Pillow / torchvision
(the third-party code that does the arithmetic) and, where the reference tree is mounted, to the reference's own
helper functions; the product's host-side geometry and coefficient tables must equal the oracle's."""
import os
import sys

import numpy as np
import pytest
import torch
from PIL import Image

import helpers
from oracle import detect_oracle as DO

CASES = [  # frame (H, W), scale, patch (w, h)
    ((120, 200), 0.75, (64, 64)),    # shrink, several patches with overlap
    ((90, 130), 1.6, (96, 80)),      # enlarge
    ((60, 70), 0.5, (64, 64)),       # scaled frame smaller than the patch: padded with 127 on both axes
    ((100, 150), 1.0, (64, 48)),     # scale 1: no resampling at all
]


def _frame(h, w, seed):
    return DO.synth_frames(1, h, w, seed=seed)[0]


@pytest.mark.parametrize("hw,scale,patch", CASES)
def test_oracle_tile_scale_equals_pillow(hw, scale, patch):
    """oracle.tile_scale == the reference's sequence executed with Pillow / torchvision themselves."""
    import torchvision.transforms.functional as TF

    h, w = hw
    frame = _frame(h, w, seed=h + w)
    img = Image.fromarray(frame)
    scaled = img.resize((int(w * scale), int(h * scale)), Image.LANCZOS)  # scale_image (ANTIALIAS == LANCZOS)
    vert_pad, horiz_pad = DO.pre_tile_padding(scaled.size[0], scaled.size[1], *patch)
    padded = TF.pad(scaled, padding=(horiz_pad, vert_pad, horiz_pad, vert_pad), fill=(127, 127, 127),
                    padding_mode="constant")
    n = DO.get_patch_spacings(padded.size[0], padded.size[1], *patch)[2]
    for idx in range(n):
        bnd = DO.patch_boundary(padded.size[0], padded.size[1], patch[0], patch[1], idx)
        want = TF.to_tensor(padded.crop(bnd))
        got, bnd2, pads, n2 = DO.tile_scale(frame, scale, patch[0], patch[1], idx)
        assert n2 == n and bnd2 == bnd and pads == (horiz_pad, vert_pad)
        assert torch.equal(torch.from_numpy(got), want), (hw, scale, patch, idx)


def test_oracle_helpers_equal_the_reference():
    ref_utils = os.path.join(helpers.REF_ROOT, "CVC-YOLOv3", "utils", "utils.py")
    if not os.path.exists(ref_utils):
        pytest.skip("reference tree not available (GPU box)")
    import importlib.util

    spec = importlib.util.spec_from_file_location("_ref_utils_for_tiles", ref_utils)
    R = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(R)
    rng = np.random.RandomState(0)
    for _ in range(40):
        pw, ph = int(rng.randint(32, 200)), int(rng.randint(32, 200))
        w, h = int(rng.randint(pw, 900)), int(rng.randint(ph, 700))
        assert DO.get_patch_spacings(w, h, pw, ph) == R.get_patch_spacings(w, h, pw, ph)
        assert DO.pre_tile_padding(w // 3, h // 3, pw, ph) == R.pre_tile_padding(w // 3, h // 3, pw, ph)
        n = R.get_patch_spacings(w, h, pw, ph)[2]
        img = Image.new("RGB", (w, h))
        for idx in (0, n - 1, n // 2):
            assert DO.patch_boundary(w, h, pw, ph, idx) == R.get_patch(img, pw, ph, idx)[1]
    # labels: the reference's chain (datasets.py:176-186,300-313) vs oracle.tile_labels
    labels_xyhw = torch.tensor([[10.0, 20.0, 30.0, 12.0], [100.0, 40.0, 50.0, 20.0], [5.0, 5.0, 8.0, 4.0]])
    scale, hp, vp, bnd, pw, ph = 0.8, 3, 0, (40.0, 10.5, 168.0, 106.5), 128, 96
    lab = R.add_class_dimension_to_labels(labels_xyhw)
    lab = R.xyhw2xyxy_corner(lab)
    lab = R.add_padding_on_each_side(R.scale_labels(lab, scale), hp, vp)
    lab = R.filter_and_offset_labels(lab, bnd)
    lab[:, 1:5] = R.xyxy2xywh(lab[:, 1:5])
    lab[:, (1, 3)] /= pw
    lab[:, (2, 4)] /= ph
    want = torch.nn.functional.pad(lab, pad=[0, 0, 0, 6 - len(lab)])
    got = DO.tile_labels(labels_xyhw.tolist(), scale, hp, vp, bnd, pw, ph, 6)
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("hw,scale,patch", CASES)
def test_product_geometry_and_tables_equal_the_oracle(hw, scale, patch):
    """b200cv.tiler computes its geometry and LANCZOS tables on the host (vectorised): same numbers as the oracle's loops."""
    from b200cv import tiler

    h, w = hw
    new_w, new_h = int(w * scale), int(h * scale)
    if new_w != w:
        for a, b in zip(tiler.lanczos_tables(w, new_w), DO.pil_lanczos_coeffs(w, new_w)):
            assert np.array_equal(a, b)
    if new_h != h:
        for a, b in zip(tiler.lanczos_tables(h, new_h), DO.pil_lanczos_coeffs(h, new_h)):
            assert np.array_equal(a, b)
    vp, hp = DO.pre_tile_padding(new_w, new_h, *patch)
    assert tiler.pre_tile_padding(new_w, new_h, *patch) == (vp, hp)
    assert tiler.get_patch_spacings(new_w + 2 * hp, new_h + 2 * vp, *patch) == \
        DO.get_patch_spacings(new_w + 2 * hp, new_h + 2 * vp, *patch)
