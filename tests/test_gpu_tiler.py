"""Tile-and-scale kernel (b200cv.tiler.TileScale) vs the oracle (pinned to Pillow): bit-exact patches, labels."""
import numpy as np
import pytest
import torch

from oracle import detect_oracle as DO

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("hw,scale,patch", [((120, 200), 0.75, (64, 64)), ((90, 130), 1.6, (96, 80)),
                                            ((60, 70), 0.5, (64, 64)), ((100, 150), 1.0, (64, 48)),
                                            ((720, 1280), 0.6, (416, 416))])
def test_tile_scale_bit_exact(hw, scale, patch):
    from b200cv.tiler import TileScale

    h, w = hw
    frames = DO.synth_frames(3, h, w, seed=7)
    ts = TileScale((h, w), scale, patch, DEV)
    rng = np.random.RandomState(1)
    idx = [0, ts.n_patches - 1, int(rng.randint(0, ts.n_patches))]
    out = ts(torch.from_numpy(frames).to(DEV), idx).cpu()
    assert out.shape == (3, 3, patch[1], patch[0])
    for b, i in enumerate(idx):
        want, bnd, pads, n = DO.tile_scale(frames[b], scale, patch[0], patch[1], i)
        assert n == ts.n_patches and bnd == ts.boundary(i) and pads == (ts.horiz_pad, ts.vert_pad)
        assert torch.equal(out[b], torch.from_numpy(want)), (hw, scale, patch, i)
    labels = [[10.0, 20.0, 30.0, 12.0], [w * 0.5, h * 0.4, h * 0.2, w * 0.1], [5.0, 5.0, 8.0, 4.0]]
    for i in idx:
        got = ts.labels(labels, i, 8)
        want = DO.tile_labels(labels, scale, ts.horiz_pad, ts.vert_pad, ts.boundary(i), patch[0], patch[1], 8)
        assert torch.allclose(got, want, rtol=1e-6, atol=1e-7)


def test_tiles_feed_the_detector(cfg_dir):
    """Patches of one 1280x720 frame straight into Darknet.forward (the reference's training input in ts mode)."""
    import helpers
    from b200cv.tiler import TileScale

    ts = TileScale((720, 1280), 0.5, (128, 128), DEV)
    frames = torch.from_numpy(DO.synth_frames(1, 720, 1280, seed=3)).to(DEV).expand(ts.n_patches, -1, -1, -1).contiguous()
    x = ts(frames, list(range(ts.n_patches)))
    model, _ = helpers.make_darknet(cfg_dir, "yolo_baseline_tiny.cfg", 128, 1)
    model = model.to(DEV).eval()
    with torch.no_grad():
        det = model(x)
    assert det.shape[0] == ts.n_patches and bool(torch.isfinite(det).all())
