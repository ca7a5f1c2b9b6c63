"""Host-side logic: cfg generation/parsing, module construction parity with the reference recipe."""
import os

import pytest
import torch

import helpers
from b200cv import cfg_gen
from oracle import yolo_oracle as YO


@pytest.mark.parametrize("kind,ref_name", [("tiny", "yolo_baseline_tiny.cfg"), ("darknet53", "yolo_baseline.cfg")])
def test_generated_cfg_equals_reference_cfg(cfg_dir, kind, ref_name):
    ref = os.path.join(helpers.REF_ROOT, "CVC-YOLOv3", "model_cfg", ref_name)
    if not os.path.exists(ref):
        pytest.skip("reference tree not available (GPU box)")
    from utils.parse_config import parse_model_config

    mine = parse_model_config(cfg_gen.write_cfg(cfg_dir, kind, 800, 800, 80))
    theirs = YO.parse_cfg(ref)
    assert len(mine) == len(theirs)
    for a, b in zip(mine[1:], theirs[1:]):
        keys = ("type", "filters", "size", "stride", "layers", "from")
        assert {k: a.get(k) for k in keys} == {k: b.get(k) for k in keys}
    for k in ("classes", "channels", "yolo_masks", "leaky_slope", "conv_activation", "build_targets_ignore_thresh",
              "conf_thresh", "nms_thresh", "iou_thresh", "start_weights_dim", "width", "height", "onnx_height"):
        assert mine[0][k] == theirs[0][k], k


def test_parse_config_matches_oracle(cfg_dir):
    from utils.parse_config import parse_model_config

    path = cfg_gen.write_cfg(cfg_dir, "tiny", 416, 416, 1)
    assert parse_model_config(path) == YO.parse_cfg(path)


@pytest.mark.parametrize("name", ["tiny_128", "full_128", "tiny_128_c80"])
def test_model_construction_matches_reference(cfg_dir, golden_yolo, name):
    g = golden_yolo["darknet"][name]
    model, _ = helpers.make_darknet(cfg_dir, g["cfg"], g["S"], g["C"])
    helpers.assert_digest(model.named_parameters(), g["digest"])
    spec = YO.NetSpec(cfg_gen.write_cfg(cfg_dir, helpers.NET_KIND[g["cfg"]], g["S"], g["S"], g["C"]))
    assert [L["filters"] for L in spec.layers if L["type"] == "convolutional"] == \
           [m[0].out_channels for d, m in zip(model.module_defs, model.module_list) if d["type"] == "convolutional"]


def test_weights_file_roundtrip(cfg_dir, tmp_path):
    model, _ = helpers.make_darknet(cfg_dir, "yolo_baseline_tiny.cfg", 128, 1)
    for b in model.buffers():
        if b.dtype == torch.float32:
            b.uniform_(0.5, 1.5)
    path = str(tmp_path / "w.weights")
    model.save_weights(path, cutoff=len(model.module_list))
    other, _ = helpers.make_darknet(cfg_dir, "yolo_baseline_tiny.cfg", 128, 1, seed=5)
    other.load_weights(path, [model.module_list[15][0].out_channels, model.module_list[22][0].out_channels])
    for (k, a), (_, b) in zip(model.state_dict().items(), other.state_dict().items()):
        if "num_batches" not in k:
            assert torch.equal(a, b), k


def _state_items(model):
    return list(model.named_parameters()) + [(k, v) for k, v in model.named_buffers() if v.dtype.is_floating_point]


def test_weights_files_written_by_the_reference(cfg_dir, tmp_path):
    """a-9: `.weights` bytes pinned to files the REFERENCE's save_weights wrote (oracle/gen_golden_weights.py):
    A = 80-class model (255-filter heads), B = the 1-class model after load_weights(A, [255, 255]) kept the first 18 of
    the 255 head filters (models.py:380-394).  Loading gives the reference's tensors; saving reproduces its bytes,
    header included (`seen` lives in header[3], models.py:344,403)."""
    import models

    gold_dir = os.path.join(os.path.dirname(__file__), "golden")
    g = torch.load(os.path.join(gold_dir, "weights_golden.pt"), weights_only=False)
    a_path, b_path = os.path.join(gold_dir, "mini_c80_ref.weights"), os.path.join(gold_dir, "mini_c1_ref.weights")
    m80 = models.Darknet(helpers.write_mini_cfg(cfg_dir, 80, g["layers"]), 2.0, 1.6, 25.0, 0.1, True)
    m80.load_weights(a_path, m80.get_start_weight_dim())
    assert int(m80.seen) == 777
    helpers.assert_digest(_state_items(m80), g["digest_c80"])
    out_a = str(tmp_path / "a.weights")
    m80.seen = 777
    m80.save_weights(out_a)
    assert open(out_a, "rb").read() == open(a_path, "rb").read()
    m1 = models.Darknet(helpers.write_mini_cfg(cfg_dir, 1, g["layers"]), 2.0, 1.6, 25.0, 0.1, True)
    m1.load_weights(a_path, m1.get_start_weight_dim())  # truncation 255 -> 18 filters per head
    assert int(m1.seen) == g["seen_after_load"] == 777
    helpers.assert_digest(_state_items(m1), g["digest_c1"])
    m1.seen = 778
    out_b = str(tmp_path / "b.weights")
    m1.save_weights(out_b)
    assert open(out_b, "rb").read() == open(b_path, "rb").read()
    # a model that never loaded a file can still save (the reference cannot: its header is a torch tensor, :277)
    fresh = models.Darknet(helpers.write_mini_cfg(cfg_dir, 1, g["layers"]), 2.0, 1.6, 25.0, 0.1, True)
    fresh.save_weights(str(tmp_path / "c.weights"))
    assert os.path.getsize(str(tmp_path / "c.weights")) == os.path.getsize(b_path)


def test_getters(cfg_dir):
    model, path = helpers.make_darknet(cfg_dir, "yolo_baseline_tiny.cfg", 416, 1)
    assert model.img_size() == (416, 416)
    assert model.get_loss_constant() == [2.0, 1.6, 25.0, 0.1]
    assert model.get_anchors() == cfg_gen.VANILLA_ANCHORS
    assert model.get_threshs() == (0.8, 0.25, 0.5)
    assert model.get_num_classes() == 1 and model.get_bw() is False
    assert model.get_onnx_name() == os.path.basename(path).split(".")[0] + "_416320.onnx"


def test_product_synthetic_recipes_equal_the_oracle_recipes():
    """bench.py and the tools draw their synthetic inputs from b200cv.synth (product side, no oracle); the parity tests
    draw theirs from the oracle modules.  Both must be the same data."""
    import numpy as np

    from b200cv import synth
    from oracle import detect_oracle as DO
    from oracle import rektnet_oracle as RO
    from oracle import yolo_oracle as YO

    assert torch.equal(synth.synth_images(2, 32, 48, seed=3), YO.synth_images(2, 32, 48, seed=3))
    assert torch.equal(synth.synth_targets(5, 16, seed=4), YO.synth_targets(5, 16, seed=4))
    for a, b in zip(synth.synth_keypoint_batch(3, seed=2), RO.synth_batch(3, seed=2)):
        assert torch.equal(a, b)
    assert np.array_equal(synth.synth_frames(2, 20, 30, seed=1), DO.synth_frames(2, 20, 30, seed=1))


def test_conv_flop_table_from_the_executor_matches_the_survey(cfg_dir):
    """bench.py counts the algorithmic conv FLOPs from the product's own layer list: SURVEY 8a-3 gives 65.86 GFLOP
    forward per image for Darknet-53 at 416x416 with C=80, 5.44 for the tiny network (C=1)."""
    import bench

    from b200cv import synth

    model, _ = helpers.make_darknet(cfg_dir, "yolo_baseline.cfg", 416, 80)
    fwd, tot = bench.conv_flops_per_image(synth.conv_layer_table(model), 416)
    assert abs(fwd / 1e9 - 65.86) < 0.01 and abs(tot / 1e9 - 197.293) < 0.01
    tiny, _ = helpers.make_darknet(cfg_dir, "yolo_baseline_tiny.cfg", 416, 1)
    fwd_t, _ = bench.conv_flops_per_image(synth.conv_layer_table(tiny), 416)
    assert abs(fwd_t / 1e9 - 5.44) < 0.01
