"""ONNX export of the two networks without the `onnx` package.

* ``darknet_to_onnx(model)`` builds the graph the reference's CVC-YOLOv3/yolo2onnx.py builds from a cfg + ``.weights``
  pair (``GraphBuilderONNX`` :268-628, ``do_everything`` :629-665): node/tensor names ``NNN_<type>`` counted over the
  cfg blocks with ``000_net`` as the input, ``Conv`` -> ``BatchNormalization`` (eps 1e-5, momentum 0.99) ->
  ``LeakyRelu``/``Relu`` per convolutional block, ``Add`` for shortcuts, ``Concat(axis=1)`` for two-input routes
  (a one-input route creates no node: the next block reads the routed tensor), ``Upsample(nearest, scales)``,
  ``MaxPool`` with the reference's pads ``[k-s, k-s, 0, 0]``; yolo blocks create no node, the graph outputs are the
  pre-YOLO convolution outputs; input ``[1, channels, onnx_height, width]``.  Weights come from the live module (what
  ``save_weights`` would write), initializers named ``NNN_convolutional_{conv_weights,conv_bias,bn_scale,...}``.
* ``keypointnet_to_onnx(model)``: the ``onnx_mode`` KeypointNet (RektNet/train_eval.py:92-99 exports it with
  ``torch.onnx.export``): stem, four residual blocks, 1x1 head; output = raw heat-map logits.
* ``do_everything(cfg_name, weights_name)``: drop-in for yolo2onnx.py's entry point (writes ``<cfg>_<w><h>.onnx``).

The files are plain ONNX protobufs (ir_version 3, opset 8 -- the ``Upsample`` with a ``scales`` attribute the reference
emits), serialised by the ~40-line encoder below; ``parse_model`` reads them back (tests run the graph with it and
compare against the B200 model, and validate the bytes with the ONNX checker bundled in torch).
"""
from __future__ import annotations

import struct

import numpy as np
import torch

IR_VERSION, OPSET = 3, 8
FLOAT, INT, STRING, FLOATS, INTS = 1, 2, 3, 6, 7  # AttributeProto.AttributeType
TENSOR_FLOAT = 1                                    # TensorProto.DataType


# ------------------------------------------------------------------------------------------ protobuf wire format
def _varint(v: int) -> bytes:
    v &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def _key(field: int, wire: int) -> bytes:
    return _varint((field << 3) | wire)


def _int(field: int, v: int) -> bytes:
    return _key(field, 0) + _varint(int(v))


def _bytes(field: int, b) -> bytes:
    b = b.encode() if isinstance(b, str) else bytes(b)
    return _key(field, 2) + _varint(len(b)) + b


def _f32(field: int, v: float) -> bytes:
    return _key(field, 5) + struct.pack("<f", float(v))


def _attr(name: str, value) -> bytes:
    """AttributeProto: name=1, f=2, i=3, s=4, floats=7, ints=8, type=20."""
    body = _bytes(1, name)
    if isinstance(value, float):
        body += _f32(2, value) + _int(20, FLOAT)
    elif isinstance(value, int):
        body += _int(3, value) + _int(20, INT)
    elif isinstance(value, str):
        body += _bytes(4, value) + _int(20, STRING)
    elif value and isinstance(value[0], float):
        body += b"".join(_f32(7, v) for v in value) + _int(20, FLOATS)
    else:
        body += b"".join(_int(8, v) for v in value) + _int(20, INTS)
    return body


def _node(op_type: str, inputs, outputs, name: str, **attrs) -> bytes:
    """NodeProto: input=1, output=2, name=3, op_type=4, attribute=5."""
    body = b"".join(_bytes(1, i) for i in inputs) + b"".join(_bytes(2, o) for o in outputs)
    body += _bytes(3, name) + _bytes(4, op_type)
    body += b"".join(_bytes(5, _attr(k, v)) for k, v in attrs.items())
    return body


def _tensor(name: str, array: np.ndarray) -> bytes:
    """TensorProto: dims=1, data_type=2, name=8, raw_data=9."""
    a = np.ascontiguousarray(array, dtype="<f4")
    return b"".join(_int(1, d) for d in a.shape) + _int(2, TENSOR_FLOAT) + _bytes(8, name) + _bytes(9, a.tobytes())


def _value_info(name: str, dims) -> bytes:
    """ValueInfoProto{name=1, type=2: TypeProto{tensor_type=1: {elem_type=1, shape=2: {dim=1: {dim_value=1}}}}}."""
    shape = b"".join(_bytes(1, _int(1, d)) for d in dims)
    return _bytes(1, name) + _bytes(2, _bytes(1, _int(1, TENSOR_FLOAT) + _bytes(2, shape)))


def _model(nodes, name, inputs, outputs, initializers, producer) -> bytes:
    """ModelProto{ir_version=1, producer_name=2, graph=7, opset_import=8}; GraphProto{node=1, name=2, initializer=5,
    input=11, output=12}."""
    graph = b"".join(_bytes(1, n) for n in nodes) + _bytes(2, name)
    graph += b"".join(_bytes(5, t) for t in initializers)
    graph += b"".join(_bytes(11, v) for v in inputs) + b"".join(_bytes(12, v) for v in outputs)
    return _int(1, IR_VERSION) + _bytes(2, producer) + _bytes(7, graph) + _bytes(8, _bytes(1, "") + _int(2, OPSET))


# ------------------------------------------------------------------------------------------ Darknet
def darknet_to_onnx(model) -> bytes:
    hp = model.hyperparams
    channels, width, onnx_height = int(hp["channels"]), int(hp["width"]), int(hp["onnx_height"])
    img_w, img_h = width, onnx_height  # DarkNetParser.get_size (:85-86): the export runs at width x onnx_height
    scales = [int(v) for v in hp["yolo_scales"].split(",")]
    slope = float(hp["leaky_slope"])
    activation = hp["conv_activation"]
    nodes, inits, graph_inputs, outputs = [], [], [_value_info("000_net", [1, channels, onnx_height, width])], []
    specs = [("000_net", channels)]  # MajorNodeSpecs of yolo2onnx.py: (output name, channels); (name, None) = no node

    def previous(target=-1):
        for name, ch in specs[target::-1]:
            if name is not None and isinstance(ch, int) and ch > 0:
                return name, ch
        raise ValueError("no previous ONNX node")

    def add_init(name, t):
        a = t.detach().cpu().float().numpy()
        inits.append(_tensor(name, a))
        graph_inputs.append(_value_info(name, list(a.shape)))  # the reference lists every initializer as an input

    head = 0
    for i, (d, m) in enumerate(zip(model.module_defs, model.module_list)):
        lname = "%03d_%s" % (i + 1, d["type"])
        kind = d["type"]
        if kind == "convolutional":
            prev, _ = previous()
            conv = m[0]
            k, s = conv.kernel_size[0], conv.stride[0]
            pad = (k - 1) // 2
            is_head = d["filters"] == "preyolo"
            ins = [prev, lname + "_conv_weights"] + ([lname + "_conv_bias"] if is_head else [])
            nodes.append(_node("Conv", ins, [lname], lname, kernel_shape=[k, k], strides=[s, s],
                               pads=[pad, pad, pad, pad], dilations=[1, 1]))
            out_name = lname
            if not is_head:
                bn = m[1]
                # initializer order of WeightLoader.load_conv_weights (:186-203): scale, bias, mean, var, conv weights
                add_init(lname + "_bn_scale", bn.weight)
                add_init(lname + "_bn_bias", bn.bias)
                add_init(lname + "_bn_mean", bn.running_mean)
                add_init(lname + "_bn_var", bn.running_var)
                nodes.append(_node("BatchNormalization",
                                   [lname] + [lname + "_bn_" + sfx for sfx in ("scale", "bias", "mean", "var")],
                                   [lname + "_bn"], lname + "_bn", epsilon=1e-5, momentum=0.99))
                out_name = lname + "_bn"
            else:
                add_init(lname + "_conv_bias", conv.bias)
            add_init(lname + "_conv_weights", conv.weight)
            if not is_head:  # the pre-YOLO convolutions are linear (yolo2onnx.py:467)
                if activation == "leaky":
                    nodes.append(_node("LeakyRelu", [out_name], [lname + "_lrelu"], lname + "_lrelu", alpha=slope))
                    out_name = lname + "_lrelu"
                elif activation == "ReLU":
                    nodes.append(_node("Relu", [out_name], [lname + "_relu"], lname + "_relu"))
                    out_name = lname + "_relu"
            else:
                # yolo2onnx.py:645 declares [filters, img_w/scale, img_h/scale] -- (W, H) order, transposed for its own
                # 800 x 320 input; the true NCHW dims are written here so that shape inference accepts the file
                outputs.append(_value_info(lname, [1, conv.out_channels, int(img_h / scales[head]),
                                                   int(img_w / scales[head])]))
                head += 1
            specs.append((out_name, conv.out_channels))
        elif kind == "shortcut":
            a, ch = previous()
            b, _ = previous(int(d["from"]))
            nodes.append(_node("Add", [a, b], [lname], lname))
            specs.append((lname, ch))
        elif kind == "route":
            idx = [int(v) for v in d["layers"].split(",")]
            if len(idx) == 1:
                # no node: drop the specs after the routed layer so that it becomes "the previous node" (:540-547)
                del specs[idx[0] + 1:]
                specs.append((None, None))
            else:
                ins, ch = [], 0
                for j in idx:
                    n_, c_ = previous(j + 1 if j > 0 else j)
                    ins.append(n_)
                    ch += c_
                nodes.append(_node("Concat", ins, [lname], lname, axis=1))
                specs.append((lname, ch))
        elif kind == "upsample":
            prev, ch = previous()
            f = float(d["stride"])
            nodes.append(_node("Upsample", [prev], [lname], lname, mode="nearest", scales=[1.0, 1.0, f, f]))
            specs.append((lname, ch))
        elif kind == "maxpool":
            prev, ch = previous()
            k, s = int(d["size"]), int(d["stride"])
            nodes.append(_node("MaxPool", [prev], [lname], lname, kernel_shape=[k, k], strides=[s, s],
                               pads=[k - s, k - s, 0, 0]))
            specs.append((lname, ch))
        else:  # yolo: "not supported, skipping ONNX node generation" (:372-376)
            specs.append((lname, None))
    return _model(nodes, "YOLO", graph_inputs, outputs, inits, "b200cv (graph of CVC-YOLOv3/yolo2onnx.py)")


def do_everything(cfg_name: str, weights_name: str, head_filters=None) -> str:
    """yolo2onnx.py:629-665: cfg + .weights -> <cfg stem>_<width><height>.onnx in the current directory.  The weights
    file must hold this cfg's own head filters (a file written by save_weights), or pass `head_filters`."""
    import os
    import sys

    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = os.path.join(here, "CVC-YOLOv3")
    if p not in sys.path:
        sys.path.insert(0, p)
    import models

    model = models.Darknet(cfg_name, 2.0, 1.6, 25.0, 0.1, True)
    dims = head_filters or [m[0].out_channels for d, m in zip(model.module_defs, model.module_list)
                            if d["type"] == "convolutional" and d["filters"] == "preyolo"]
    model.load_weights(weights_name, dims)
    out = cfg_name.split("/")[-1].split(".")[0] + "_" + str(model.img_width) + str(model.img_height) + ".onnx"
    with open(out, "wb") as f:
        f.write(darknet_to_onnx(model))
    return out


# ------------------------------------------------------------------------------------------ KeypointNet
def keypointnet_to_onnx(model, image_size=None) -> bytes:
    """The `onnx_mode` network of RektNet/keypoint_net.py:58-66 (what train_eval.py:92-99 exports): heat-map logits."""
    h, w = image_size or model.image_size
    nodes, inits, graph_inputs = [], [], [_value_info("input", [1, 3, h, w])]

    def add(name, t):
        a = t.detach().cpu().float().numpy()
        inits.append(_tensor(name, a))
        graph_inputs.append(_value_info(name, list(a.shape)))  # IR version 3: initializers are graph inputs too
        return name

    def conv(x, mod, name):
        k, p_, dl = mod.kernel_size[0], mod.padding[0], mod.dilation[0]
        nodes.append(_node("Conv", [x, add(name + ".weight", mod.weight), add(name + ".bias", mod.bias)], [name + "_out"],
                           name, kernel_shape=[k, k], strides=[1, 1], pads=[p_, p_, p_, p_], dilations=[dl, dl]))
        return name + "_out"

    def bn(x, mod, name):
        ins = [x, add(name + ".weight", mod.weight), add(name + ".bias", mod.bias),
               add(name + ".running_mean", mod.running_mean), add(name + ".running_var", mod.running_var)]
        nodes.append(_node("BatchNormalization", ins, [name + "_out"], name, epsilon=float(mod.eps), momentum=0.9))
        return name + "_out"

    def relu(x, name):
        nodes.append(_node("Relu", [x], [name], name))
        return name

    x = relu(bn(conv("input", model.conv, "conv"), model.bn, "bn"), "relu")
    for r in ("res1", "res2", "res3", "res4"):
        blk = getattr(model, r)
        c1 = relu(bn(conv(x, blk.conv1, r + ".conv1"), blk.bn1, r + ".bn1"), r + ".relu1")
        c2 = bn(conv(c1, blk.conv2, r + ".conv2"), blk.bn2, r + ".bn2")
        sc = bn(conv(x, blk.shortcut_conv, r + ".shortcut_conv"), blk.shortcut_bn, r + ".shortcut_bn")
        nodes.append(_node("Add", [sc, c2], [r + ".sum"], r + ".sum"))
        x = relu(r + ".sum", r + ".relu2")
    out = conv(x, model.out, "out")
    return _model(nodes, "KeypointNet", graph_inputs, [_value_info(out, [1, model.out.out_channels, h, w])], inits,
                  "b200cv")


# ------------------------------------------------------------------------------------------ reading a file back
def _fields(buf: bytes):
    i, n = 0, len(buf)
    while i < n:
        key = shift = 0
        while True:
            b = buf[i]
            i += 1
            key |= (b & 0x7F) << shift
            shift += 7
            if not b & 0x80:
                break
        field, wire = key >> 3, key & 7
        if wire == 0:
            v = shift = 0
            while True:
                b = buf[i]
                i += 1
                v |= (b & 0x7F) << shift
                shift += 7
                if not b & 0x80:
                    break
            yield field, v
        elif wire == 2:
            ln = shift = 0
            while True:
                b = buf[i]
                i += 1
                ln |= (b & 0x7F) << shift
                shift += 7
                if not b & 0x80:
                    break
            yield field, buf[i:i + ln]
            i += ln
        elif wire == 5:
            yield field, struct.unpack("<f", buf[i:i + 4])[0]
            i += 4
        else:
            raise ValueError(f"unsupported wire type {wire}")


def parse_model(data: bytes) -> dict:
    """{'ir_version', 'opset', 'nodes': [{'op','name','inputs','outputs','attrs'}], 'initializers': {name: ndarray},
    'inputs': {name: dims}, 'outputs': {name: dims}} of a file written by this module (or any fp32 raw_data model)."""
    out = {"nodes": [], "initializers": {}, "inputs": {}, "outputs": {}}

    def value_info(b):
        name, dims = None, []
        for f, v in _fields(b):
            if f == 1:
                name = v.decode()
            elif f == 2:
                for f2, t in _fields(v):
                    if f2 == 1:
                        for f3, sh in _fields(t):
                            if f3 == 2:
                                dims = [next(val for ff, val in _fields(dm) if ff == 1) for fd, dm in _fields(sh) if fd == 1]
        return name, dims

    for f, v in _fields(data):
        if f == 1:
            out["ir_version"] = v
        elif f == 8:
            out["opset"] = dict(_fields(v)).get(2)
        elif f == 7:
            for gf, gv in _fields(v):
                if gf == 1:
                    node = {"inputs": [], "outputs": [], "attrs": {}}
                    for nf, nv in _fields(gv):
                        if nf == 1:
                            node["inputs"].append(nv.decode())
                        elif nf == 2:
                            node["outputs"].append(nv.decode())
                        elif nf == 3:
                            node["name"] = nv.decode()
                        elif nf == 4:
                            node["op"] = nv.decode()
                        elif nf == 5:
                            a = {"floats": [], "ints": []}
                            for af, av in _fields(nv):
                                if af == 1:
                                    a["name"] = av.decode()
                                elif af in (2, 3):
                                    a["v"] = av
                                elif af == 4:
                                    a["v"] = av.decode()
                                elif af == 7:
                                    a["floats"].append(av)
                                elif af == 8:
                                    a["ints"].append(av)
                                elif af == 20:
                                    a["type"] = av
                            node["attrs"][a["name"]] = a["floats"] if a["type"] == FLOATS else (
                                a["ints"] if a["type"] == INTS else a["v"])
                    out["nodes"].append(node)
                elif gf == 5:
                    dims, name, raw = [], None, b""
                    for tf, tv in _fields(gv):
                        if tf == 1:
                            dims.append(tv)
                        elif tf == 8:
                            name = tv.decode()
                        elif tf == 9:
                            raw = tv
                    out["initializers"][name] = np.frombuffer(raw, dtype="<f4").reshape(dims)
                elif gf in (11, 12):
                    name, dims = value_info(gv)
                    out["inputs" if gf == 11 else "outputs"][name] = dims
    return out


def run_graph(parsed: dict, x: torch.Tensor) -> dict:
    """Evaluate a parsed graph with torch CPU ops (test aid for the seven operators the exporters emit).  Returns every
    graph output."""
    import torch.nn.functional as F

    env = {k: torch.from_numpy(np.array(v)) for k, v in parsed["initializers"].items()}
    env[next(n for n in parsed["inputs"] if n not in parsed["initializers"])] = x
    for nd in parsed["nodes"]:
        a, ins = nd["attrs"], [env[i] for i in nd["inputs"]]
        op = nd["op"]
        if op == "Conv":
            y = F.conv2d(ins[0], ins[1], ins[2] if len(ins) > 2 else None, stride=a["strides"], padding=a["pads"][:2],
                         dilation=a["dilations"])
        elif op == "BatchNormalization":
            y = F.batch_norm(ins[0], ins[3], ins[4], ins[1], ins[2], False, 0.0, a["epsilon"])
        elif op == "LeakyRelu":
            y = F.leaky_relu(ins[0], a["alpha"])
        elif op == "Relu":
            y = F.relu(ins[0])
        elif op == "Add":
            y = ins[0] + ins[1]
        elif op == "Concat":
            y = torch.cat(ins, a["axis"])
        elif op == "Upsample":
            y = F.interpolate(ins[0], scale_factor=a["scales"][2], mode="nearest")
        elif op == "MaxPool":
            p = a["pads"]  # [top, left, bottom, right]; ONNX pads max-pools with -inf
            y = F.max_pool2d(F.pad(ins[0], (p[1], p[3], p[0], p[2]), value=float("-inf")), a["kernel_shape"], a["strides"])
        else:
            raise ValueError(f"operator {op} not handled")
        env[nd["outputs"][0]] = y
    return {k: env[k] for k in parsed["outputs"]}
