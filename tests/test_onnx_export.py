"""ONNX export without the onnx package: the bytes pass the ONNX checker bundled in torch, the graph has
the structure CVC-YOLOv3/yolo2onnx.py builds (names, operators, attributes, initializer order), and evaluating the
parsed graph reproduces the oracle's eval-mode forward (pinned to the reference) on the same weights."""
import os

import pytest
import torch

import helpers
from b200cv import onnx_export as OX
from oracle import rektnet_oracle as RO
from oracle import yolo_oracle as YO


def _check(data: bytes):
    torch._C._check_onnx_proto(data)  # onnx::checker::check_model on the serialized ModelProto


def _randomize(model, seed=3):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for k, b in model.named_buffers():
            if k.endswith("running_mean"):
                b.copy_(torch.randn(b.shape, generator=g) * 0.1)
            elif k.endswith("running_var"):
                b.copy_(torch.rand(b.shape, generator=g) + 0.5)


def test_darknet53_graph_structure_and_numerics(cfg_dir):
    model, path = helpers.make_darknet(cfg_dir, "yolo_baseline.cfg", 64, 2)
    _randomize(model)
    data = OX.darknet_to_onnx(model)
    _check(data)
    g = OX.parse_model(data)
    assert g["ir_version"] == OX.IR_VERSION and g["opset"] == OX.OPSET
    ops = [n["op"] for n in g["nodes"]]
    assert (ops.count("Conv"), ops.count("BatchNormalization"), ops.count("LeakyRelu"), ops.count("Add"),
            ops.count("Concat"), ops.count("Upsample")) == (75, 72, 72, 23, 2, 2)
    # the well-known names of the reference's graph: input 000_net, outputs = the three pre-YOLO convolutions
    assert list(g["outputs"]) == ["082_convolutional", "094_convolutional", "106_convolutional"]
    assert g["inputs"]["000_net"] == [1, 3, int(model.hyperparams["onnx_height"]), 64]
    assert g["outputs"]["082_convolutional"] == [1, 21, 10, 2]  # [1, 3*(5+C), onnx_height/32, width/32]
    first = g["nodes"][0]
    assert first["name"] == "001_convolutional" and first["inputs"] == ["000_net", "001_convolutional_conv_weights"]
    assert first["attrs"] == {"kernel_shape": [3, 3], "strides": [1, 1], "pads": [1, 1, 1, 1], "dilations": [1, 1]}
    bn = g["nodes"][1]
    assert bn["op"] == "BatchNormalization" and abs(bn["attrs"]["epsilon"] - 1e-5) < 1e-12
    assert abs(bn["attrs"]["momentum"] - 0.99) < 1e-7
    assert bn["inputs"] == ["001_convolutional"] + ["001_convolutional_bn_" + s for s in ("scale", "bias", "mean", "var")]
    # initializer order of WeightLoader.load_conv_weights (yolo2onnx.py:186-207)
    assert list(g["initializers"])[:5] == ["001_convolutional_bn_scale", "001_convolutional_bn_bias",
                                           "001_convolutional_bn_mean", "001_convolutional_bn_var",
                                           "001_convolutional_conv_weights"]
    # the one-input routes create no node: block 84 (0-based 83 = "route -4") hands block 80's output to conv 85
    conv85 = next(n for n in g["nodes"] if n["name"] == "085_convolutional")
    assert conv85["inputs"][0] == "080_convolutional_lrelu"
    cat = next(n for n in g["nodes"] if n["name"] == "087_route")
    assert cat["inputs"] == ["086_upsample", "062_shortcut"] and cat["attrs"]["axis"] == 1
    # numerics: raw head outputs, decoded by the oracle's yolo_layer, equal the oracle's eval forward
    x = YO.synth_images(1, 64, 64, seed=2)  # (the graph itself is size-agnostic: evaluated at the training size here)
    heads = OX.run_graph(g, x)
    spec = YO.NetSpec(path)
    params = {k: v.detach() for k, v in model.named_parameters()}
    buffers = dict(model.named_buffers())
    with torch.no_grad():
        want = YO.darknet_forward(spec, params, buffers, x, None, training=False)
    yolo_layers = [L for L in spec.layers if L["type"] == "yolo"]
    got = torch.cat([YO.yolo_layer(heads[n], None, L["anchors"], spec.num_classes, spec.height, spec.ignore_thresh, 2.0,
                                   1.6, 0.1, 25.0) for n, L in zip(g["outputs"], yolo_layers)], 1)
    assert torch.allclose(got, want, rtol=1e-4, atol=1e-5)


def test_tiny_graph_maxpool_quirk_and_checker(cfg_dir):
    model, _ = helpers.make_darknet(cfg_dir, "yolo_baseline_tiny.cfg", 64, 1)
    data = OX.darknet_to_onnx(model)
    _check(data)
    g = OX.parse_model(data)
    pools = [n for n in g["nodes"] if n["op"] == "MaxPool"]
    assert len(pools) == 6
    assert pools[0]["attrs"] == {"kernel_shape": [2, 2], "strides": [2, 2], "pads": [0, 0, 0, 0]}
    assert pools[5]["attrs"]["strides"] == [1, 1] and pools[5]["attrs"]["pads"] == [1, 1, 0, 0]  # yolo2onnx.py:606-610
    assert list(g["outputs"]) == ["016_convolutional", "023_convolutional"]


def test_do_everything_from_cfg_and_weights(cfg_dir, tmp_path, monkeypatch):
    """yolo2onnx.py's entry point: cfg + .weights (the reference-written golden) -> <cfg>_<w><h>.onnx."""
    gold_dir = os.path.join(os.path.dirname(__file__), "golden")
    g = torch.load(os.path.join(gold_dir, "weights_golden.pt"), weights_only=False)
    cfg = helpers.write_mini_cfg(cfg_dir, 1, g["layers"])
    monkeypatch.chdir(tmp_path)
    out = OX.do_everything(cfg, os.path.join(gold_dir, "mini_c1_ref.weights"))
    assert out == "mini_c1_6464.onnx" and os.path.exists(tmp_path / out)
    data = open(tmp_path / out, "rb").read()
    _check(data)
    parsed = OX.parse_model(data)
    import models

    ref = models.Darknet(cfg, 2.0, 1.6, 25.0, 0.1, True)
    ref.load_weights(os.path.join(gold_dir, "mini_c1_ref.weights"), [18, 18])
    w = parsed["initializers"]["001_convolutional_conv_weights"]
    assert torch.equal(torch.from_numpy(w.copy()), ref.module_list[0][0].weight.detach())


def test_keypointnet_graph(golden_rekt):
    import keypoint_net

    torch.manual_seed(17)
    net = keypoint_net.KeypointNet(onnx_mode=True)
    _randomize(net, seed=5)
    data = OX.keypointnet_to_onnx(net)
    _check(data)
    g = OX.parse_model(data)
    ops = [n["op"] for n in g["nodes"]]
    assert (ops.count("Conv"), ops.count("BatchNormalization"), ops.count("Relu"), ops.count("Add")) == (14, 13, 9, 4)
    assert g["inputs"]["input"] == [1, 3, 80, 80] and list(g["outputs"].values()) == [[1, 7, 80, 80]]
    dil = next(n for n in g["nodes"] if n["name"] == "res1.conv1")
    assert dil["attrs"]["dilations"] == [2, 2] and dil["attrs"]["pads"] == [2, 2, 2, 2]
    x, _, _ = RO.synth_batch(1, seed=1)
    params = {k: v.detach() for k, v in net.named_parameters()}
    buffers = dict(net.named_buffers())
    with torch.no_grad():
        want = RO.keypointnet_forward(params, buffers, x, training=False, onnx_mode=True)
    got = list(OX.run_graph(g, x).values())[0]
    assert torch.allclose(got, want, rtol=1e-4, atol=1e-5)
