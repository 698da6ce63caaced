"""Minimal ONNX (protobuf) WRITER for tests: an insightface-iresnet state dict -> the graph torch.onnx.export would
emit for it, in the two flavours seen in the wild - BatchNormalization kept as nodes with torch parameter names, or
every BN that follows a Conv folded into it by the exporter (Conv gets a bias, initializers get generated numeric
names).  No `onnx` package needed.  Test infrastructure only (the product reader is facerecognitionpipeline_b200/
onnx_import.py; reader and writer share no code)."""
import struct

import numpy as np


def _vi(n):
    out = bytearray()
    n &= (1 << 64) - 1
    while True:
        b = n & 0x7F
        n >>= 7
        out.append(b | (0x80 if n else 0))
        if not n:
            return bytes(out)


def _ld(field, payload):
    return _vi((field << 3) | 2) + _vi(len(payload)) + payload


def _iv(field, v):
    return _vi(field << 3) + _vi(v)


def _tensor(name, arr, use_raw=True):
    arr = np.ascontiguousarray(arr)
    dt = {np.dtype(np.float32): 1, np.dtype(np.float64): 11, np.dtype(np.float16): 10, np.dtype(np.int64): 7}[arr.dtype]
    out = b"".join(_iv(1, int(d)) for d in arr.shape) + _iv(2, dt) + _ld(8, name.encode())
    if use_raw or dt != 1:
        out += _ld(9, arr.tobytes())
    else:
        out += _ld(4, arr.astype("<f4").tobytes())       # packed float_data
    return out


def _attr(name, v):
    out = _ld(1, name.encode())
    if isinstance(v, float):
        out += _vi((2 << 3) | 5) + struct.pack("<f", v) + _iv(20, 1)
    elif isinstance(v, int):
        out += _iv(3, v) + _iv(20, 2)
    else:
        out += b"".join(_iv(8, int(x)) for x in v) + _iv(20, 7)
    return out


def _node(op, inputs, outputs, name="", **attrs):
    out = b"".join(_ld(1, i.encode()) for i in inputs) + b"".join(_ld(2, o.encode()) for o in outputs)
    out += _ld(3, name.encode()) + _ld(4, op.encode())
    out += b"".join(_ld(5, _attr(k, v)) for k, v in attrs.items())
    return out


class _Graph:
    def __init__(self, numeric_names):
        self.nodes, self.inits, self.n, self.numeric = [], [], 0, numeric_names

    def t(self):
        self.n += 1
        return f"t{self.n}"

    def init(self, name, arr, raw=True):
        if self.numeric:
            self.n += 1
            name = str(1000 + self.n)
        self.inits.append(_tensor(name, arr, raw))
        return name

    def node(self, op, inputs, **attrs):
        out = self.t()
        self.nodes.append(_node(op, inputs, [out], name=f"{op}_{len(self.nodes)}", **attrs))
        return out


def write_iresnet_onnx(path, sd, units, fold_conv_bn=False, gemm=True, eps=1e-5, batch="N"):
    """sd: insightface iresnet state dict (torch tensors); units: e.g. [3, 4, 14, 3]."""
    g = _Graph(numeric_names=fold_conv_bn)
    f32 = lambda k: sd[k].detach().cpu().numpy().astype(np.float32)

    def bn_affine(prefix):
        a = f32(prefix + ".weight").astype(np.float64) / np.sqrt(f32(prefix + ".running_var").astype(np.float64) + eps)
        return a, f32(prefix + ".bias").astype(np.float64) - f32(prefix + ".running_mean").astype(np.float64) * a

    def bn(x, prefix):
        return g.node("BatchNormalization", [x] + [g.init(prefix + s, f32(prefix + s)) for s in (".weight", ".bias", ".running_mean", ".running_var")],
                      epsilon=float(eps), momentum=0.9)

    def conv_bn(x, wkey, bnprefix, k, stride):
        W = f32(wkey)
        pad = [1, 1, 1, 1] if k == 3 else [0, 0, 0, 0]
        if fold_conv_bn:                      # what the exporter's eval-mode fusion produces
            a, b = bn_affine(bnprefix)
            Wf = (W.astype(np.float64) * a[:, None, None, None]).astype(np.float32)
            return g.node("Conv", [x, g.init(wkey, Wf), g.init(wkey + ".b", b.astype(np.float32))], dilations=[1, 1], group=1,
                          kernel_shape=[k, k], pads=pad, strides=[stride, stride])
        y = g.node("Conv", [x, g.init(wkey, W, raw=False)], dilations=[1, 1], group=1, kernel_shape=[k, k], pads=pad, strides=[stride, stride])
        return bn(y, bnprefix)

    def prelu(x, key):
        return g.node("PRelu", [x, g.init(key, f32(key).reshape(-1, 1, 1))])

    x = conv_bn("input.1", "conv1.weight", "bn1", 3, 1)
    x = prelu(x, "prelu.weight")
    for stage, n in enumerate(units):
        for j in range(n):
            p = f"layer{stage + 1}.{j}."
            stride = 2 if j == 0 else 1
            y = bn(x, p + "bn1")
            y = conv_bn(y, p + "conv1.weight", p + "bn2", 3, 1)
            y = prelu(y, p + "prelu.weight")
            y = conv_bn(y, p + "conv2.weight", p + "bn3", 3, stride)
            sc = conv_bn(x, p + "downsample.0.weight", p + "downsample.1", 1, stride) if j == 0 else x
            x = g.node("Add", [y, sc])
    x = bn(x, "bn2")
    x = g.node("Identity", [x])                                         # Dropout in eval mode exports as Identity / is dropped
    x = g.node("Flatten", [x], axis=1)
    if gemm:
        x = g.node("Gemm", [x, g.init("fc.weight", f32("fc.weight")), g.init("fc.bias", f32("fc.bias"))], alpha=1.0, beta=1.0, transB=1)
    else:
        x = g.node("MatMul", [x, g.init("fc.weight.T", np.ascontiguousarray(f32("fc.weight").T))])
        x = g.node("Add", [x, g.init("fc.bias", f32("fc.bias"))])
    out = bn(x, "features")
    graph = b"".join(_ld(1, n) for n in g.nodes) + _ld(2, b"iresnet") + b"".join(_ld(5, t) for t in g.inits)
    # ValueInfoProto { name, type { tensor_type { elem_type = FLOAT, shape { dim... } } } } as every real export carries
    def vinfo(name, dims):
        shape = b"".join(_ld(1, _iv(1, int(d)) if isinstance(d, int) else _ld(2, d.encode())) for d in dims)
        return _ld(1, name.encode()) + _ld(2, _ld(1, _iv(1, 1) + _ld(2, shape)))
    graph += _ld(11, vinfo("input.1", [batch, 3, 112, 112])) + _ld(12, vinfo(out, [batch, 512]))
    model = _iv(1, 8) + _ld(2, b"frb200-test-writer") + _ld(8, _ld(1, b"") + _iv(2, 13)) + _ld(7, graph)
    with open(path, "wb") as f:
        f.write(model)


def write_adaface_onnx(path, sd, arch, eps=1e-5, batch="N"):
    """The AdaFace IR backbone (mk-minchul/AdaFace net.py as restated by oracle/backbone.py: `input_layer`, `body.N`
    with BN -> conv3x3 -> BN -> PReLU -> conv3x3(stride) -> BN and a MaxPool(1, stride) or conv1x1 + BN shortcut,
    `output_layer`) as an ONNX graph, so that an independent ONNX engine can be run against the oracle's forward.
    Output = the 512 features BEFORE the L2 normalisation."""
    from oracle import backbone as ob
    g = _Graph(numeric_names=False)
    f32 = lambda k: sd[k].detach().cpu().numpy().astype(np.float32)

    def bn(x, prefix, c):
        w = f32(prefix + ".weight") if prefix + ".weight" in sd else np.ones(c, np.float32)
        b = f32(prefix + ".bias") if prefix + ".bias" in sd else np.zeros(c, np.float32)
        ins = [g.init(prefix + ".weight", w), g.init(prefix + ".bias", b),
               g.init(prefix + ".running_mean", f32(prefix + ".running_mean")), g.init(prefix + ".running_var", f32(prefix + ".running_var"))]
        return g.node("BatchNormalization", [x] + ins, epsilon=float(eps), momentum=0.9)

    def conv(x, wkey, k, stride):
        pad = [1, 1, 1, 1] if k == 3 else [0, 0, 0, 0]
        return g.node("Conv", [x, g.init(wkey, f32(wkey))], dilations=[1, 1], group=1, kernel_shape=[k, k], pads=pad, strides=[stride, stride])

    def prelu(x, key):
        return g.node("PRelu", [x, g.init(key, f32(key).reshape(-1, 1, 1))])

    hk = ob.head_keys("adaface")
    x = prelu(bn(conv("input.1", hk["stem_conv"], 3, 1), hk["stem_bn"], 64), hk["stem_prelu"])
    for keys, in_c, d, s, has_sc in ob.unit_key_list(arch, "adaface"):
        if has_sc:
            sc = bn(conv(x, keys["sc_conv"], 1, s), keys["sc_bn"], d)
        else:
            sc = g.node("MaxPool", [x], kernel_shape=[1, 1], pads=[0, 0, 0, 0], strides=[s, s])
        r = conv(bn(x, keys["bn1"], in_c), keys["conv1"], 3, 1)
        r = prelu(bn(r, keys["bn2"], d), keys["prelu"])
        r = bn(conv(r, keys["conv2"], 3, s), keys["bn3"], d)
        x = g.node("Add", [r, sc])
    x = g.node("Flatten", [bn(x, hk["out_bn"], 512)], axis=1)
    x = g.node("Gemm", [x, g.init(hk["fc_w"], f32(hk["fc_w"])), g.init(hk["fc_b"], f32(hk["fc_b"]))], alpha=1.0, beta=1.0, transB=1)
    out = bn(x, hk["feat_bn"], 512)

    def vinfo(name, dims):
        shape = b"".join(_ld(1, _iv(1, int(d)) if isinstance(d, int) else _ld(2, d.encode())) for d in dims)
        return _ld(1, name.encode()) + _ld(2, _ld(1, _iv(1, 1) + _ld(2, shape)))
    graph = b"".join(_ld(1, n) for n in g.nodes) + _ld(2, b"adaface_ir") + b"".join(_ld(5, t) for t in g.inits)
    graph += _ld(11, vinfo("input.1", [batch, 3, 112, 112])) + _ld(12, vinfo(out, [batch, 512]))
    model = _iv(1, 8) + _ld(2, b"frb200-test-writer") + _ld(8, _ld(1, b"") + _iv(2, 13)) + _ld(7, graph)
    with open(path, "wb") as f:
        f.write(model)
