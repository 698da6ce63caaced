"""ONNX weight import for the ArcFace branch (reference face_embedder.py:64-88: `ort.InferenceSession(arcface_*.onnx)`).

The reference runs insightface's `iresnet50/100` as an ONNX export through onnxruntime.  Here the graph is not
executed: its initializers are read and mapped, by walking the dataflow, onto the iresnet state-dict layout that
`weights.build_program(sd, arch, "iresnet")` folds into the device program.  No `onnx` / `onnxruntime` package is
needed (neither exists in this image): ONNX is protobuf, and the handful of messages involved are decoded by the
~100-line wire-format reader below.

Handled export variants:
  * BatchNormalization kept as nodes (scale / B / mean / var initializers), or folded into the preceding Conv by the
    exporter (Conv carries a bias, the BN node is gone) - the folded case becomes an identity BatchNorm whose shift is
    the conv bias, which `weights.build_program` folds back to the same numbers;
  * initializer names: anything (torch parameter names or exporter-generated numbers) - only graph structure is used;
  * Gemm (transB 0/1, alpha/beta) or MatMul + Add for the 25088 -> 512 layer; Flatten or Reshape before it;
    Dropout / Identity nodes are skipped; tensors stored as raw_data or float_data, fp32 / fp16 / fp64.
Anything else (a preprocessing prologue, another backbone) raises ValueError naming the node it stopped at.
"""
from __future__ import annotations

import struct
from typing import Dict, List, Tuple

import numpy as np

# ------------------------------------------------------------------------------------------------ protobuf wire format


def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    val, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        val |= (b & 0x7F) << shift
        if not b & 0x80:
            return val, pos
        shift += 7
        if shift > 70:
            raise ValueError("malformed varint")


def _fields(buf: bytes):
    """Yield (field number, wire type, value) of one message; value = int for varint / fixed, memoryview for bytes."""
    pos, end = 0, len(buf)
    view = memoryview(buf)
    while pos < end:
        key, pos = _varint(buf, pos)
        num, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            n, pos = _varint(buf, pos)
            if pos + n > end:
                raise ValueError("truncated length-delimited field")
            v = view[pos:pos + n]
            pos += n
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield num, wt, v


def _signed(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


def _packed_ints(v) -> List[int]:
    out, pos, raw = [], 0, bytes(v)
    while pos < len(raw):
        x, pos = _varint(raw, pos)
        out.append(_signed(x))
    return out


_DTYPES = {1: np.float32, 10: np.float16, 11: np.float64, 7: np.int64, 6: np.int32}


def _tensor(buf) -> Tuple[str, np.ndarray]:
    """TensorProto -> (name, ndarray)."""
    dims: List[int] = []
    dtype, name, raw, floats, int64s, doubles, external = 1, "", None, [], [], [], False
    for num, wt, v in _fields(bytes(buf)):
        if num == 1:
            dims += _packed_ints(v) if wt == 2 else [_signed(v)]
        elif num == 2:
            dtype = v
        elif num == 4:
            floats.append(np.frombuffer(bytes(v), "<f4") if wt == 2 else np.array([struct.unpack("<f", struct.pack("<I", v))[0]], np.float32))
        elif num == 7:
            int64s += _packed_ints(v) if wt == 2 else [_signed(v)]
        elif num == 10:
            doubles.append(np.frombuffer(bytes(v), "<f8") if wt == 2 else np.array([struct.unpack("<d", struct.pack("<Q", v))[0]]))
        elif num == 8:
            name = bytes(v).decode("utf-8")
        elif num == 9:
            raw = bytes(v)
        elif num in (13, 14) and (num == 13 or v == 1):
            external = True
    if external:
        raise ValueError(f"initializer {name!r} uses external data files; re-export the model with embedded weights")
    if dtype not in _DTYPES:
        raise ValueError(f"initializer {name!r}: unsupported ONNX data type {dtype}")
    if raw is not None:
        arr = np.frombuffer(raw, np.dtype(_DTYPES[dtype]).newbyteorder("<"))
    elif floats:
        arr = np.concatenate(floats)
    elif doubles:
        arr = np.concatenate(doubles)
    else:
        arr = np.array(int64s, dtype=np.int64)
    return name, np.array(arr).reshape(dims if dims else ())


class _Node:
    __slots__ = ("op", "name", "inputs", "outputs", "attrs")

    def __init__(self):
        self.op, self.name, self.inputs, self.outputs, self.attrs = "", "", [], [], {}


def _node(buf) -> _Node:
    n = _Node()
    for num, wt, v in _fields(bytes(buf)):
        if num == 1:
            n.inputs.append(bytes(v).decode("utf-8"))
        elif num == 2:
            n.outputs.append(bytes(v).decode("utf-8"))
        elif num == 3:
            n.name = bytes(v).decode("utf-8")
        elif num == 4:
            n.op = bytes(v).decode("utf-8")
        elif num == 5:
            aname, val = "", None
            for an, awt, av in _fields(bytes(v)):
                if an == 1:
                    aname = bytes(av).decode("utf-8")
                elif an == 2:
                    val = struct.unpack("<f", struct.pack("<I", av))[0]
                elif an == 3:
                    val = _signed(av)
                elif an == 8:
                    val = (val if isinstance(val, list) else []) + (_packed_ints(av) if awt == 2 else [_signed(av)])
                elif an == 5:
                    val = _tensor(av)[1]
            n.attrs[aname] = val
    return n


def read_onnx(path: str):
    """-> (nodes in graph order, {initializer name: ndarray}, [graph input names that are not initializers])."""
    data = open(path, "rb").read()
    graph = None
    for num, wt, v in _fields(data):
        if num == 7 and wt == 2:
            graph = bytes(v)
    if graph is None:
        raise ValueError(f"{path}: no GraphProto found (not an ONNX model?)")
    nodes, inits, inputs = [], {}, []
    for num, wt, v in _fields(graph):
        if num == 1:
            nodes.append(_node(v))
        elif num == 5:
            name, arr = _tensor(v)
            inits[name] = arr
        elif num == 11:
            for vn, _, vv in _fields(bytes(v)):
                if vn == 1:
                    inputs.append(bytes(vv).decode("utf-8"))
    # Constant nodes are initializers in disguise
    for n in nodes:
        if n.op == "Constant" and "value" in n.attrs and n.outputs:
            inits[n.outputs[0]] = n.attrs["value"]
    return nodes, inits, [i for i in inputs if i not in inits]


# ------------------------------------------------------------------------------------------------ graph -> state dict
_SKIP = ("Dropout", "Identity", "Cast")


class _Walker:
    def __init__(self, nodes, inits):
        self.nodes = [n for n in nodes if n.op != "Constant"]
        self.inits = inits
        self.consumers: Dict[str, List[_Node]] = {}
        for n in self.nodes:
            for i in n.inputs:
                if i not in inits:
                    self.consumers.setdefault(i, []).append(n)
        self.producer = {o: n for n in self.nodes for o in n.outputs}

    def users(self, t: str) -> List[_Node]:
        out = []
        for n in self.consumers.get(t, []):
            if n.op in _SKIP:
                out += self.users(n.outputs[0])
            else:
                out.append(n)
        return out

    def only_user(self, t: str, what: str) -> _Node:
        u = self.users(t)
        if len(u) != 1:
            raise ValueError(f"ONNX import: expected exactly one consumer of {t!r} ({what}), found {[n.op for n in u]}")
        return u[0]

    def source(self, t: str) -> str:
        """Skip pass-through nodes backwards."""
        n = self.producer.get(t)
        while n is not None and n.op in _SKIP:
            t = n.inputs[0]
            n = self.producer.get(t)
        return t

    def w(self, name: str) -> np.ndarray:
        if name not in self.inits:
            raise ValueError(f"ONNX import: {name!r} is not an initializer (weights computed inside the graph are not supported)")
        return np.asarray(self.inits[name])


def _put_bn(sd, prefix, scale, shift, mean, var):
    import torch
    sd[prefix + ".weight"] = torch.from_numpy(np.asarray(scale, np.float64).copy())
    sd[prefix + ".bias"] = torch.from_numpy(np.asarray(shift, np.float64).copy())
    sd[prefix + ".running_mean"] = torch.from_numpy(np.asarray(mean, np.float64).copy())
    sd[prefix + ".running_var"] = torch.from_numpy(np.asarray(var, np.float64).copy())


def _identity_bn(sd, prefix, c, shift=None, eps=1e-5):
    """BatchNorm that only adds `shift` (a conv bias left behind by exporter-side folding): with
    running_var = 1 - eps the scale 1/sqrt(var + eps) is 1."""
    _put_bn(sd, prefix, np.ones(c), np.zeros(c) if shift is None else shift, np.zeros(c), np.full(c, 1.0 - eps))


def _take_bn(wk: _Walker, node: _Node, sd, prefix, eps_expected=1e-5):
    scale, shift, mean, var = (wk.w(node.inputs[i]) for i in (1, 2, 3, 4))
    eps = float(node.attrs.get("epsilon", 1e-5))
    # weights.build_program folds with eps = 1e-5 (the attribute is an f32: 9.99999975e-06 IS 1e-5); any other epsilon
    # is absorbed into the variance
    delta = 0.0 if abs(eps - eps_expected) < 1e-9 else eps - eps_expected
    _put_bn(sd, prefix, scale, shift, mean, np.asarray(var, np.float64) + delta)


def _conv_then_bn(wk: _Walker, conv: _Node, sd, conv_key, bn_prefix, expect_k, expect_stride) -> str:
    """Conv [+ BatchNormalization] -> tensor name after them; records the conv weight and the BN (identity when the
    exporter folded it into the conv)."""
    import torch
    if conv.op != "Conv":
        raise ValueError(f"ONNX import: expected Conv for {conv_key}, found {conv.op} ({conv.name})")
    W = wk.w(conv.inputs[1]).astype(np.float32)
    ks = conv.attrs.get("kernel_shape", list(W.shape[2:]))
    st = conv.attrs.get("strides", [1, 1])
    if list(ks) != [expect_k, expect_k] or list(st) != [expect_stride, expect_stride] or int(conv.attrs.get("group", 1)) != 1:
        raise ValueError(f"ONNX import: {conv_key}: kernel {ks} stride {st} group {conv.attrs.get('group', 1)}; "
                         f"expected {expect_k}x{expect_k} stride {expect_stride} (not an insightface iresnet?)")
    sd[conv_key] = torch.from_numpy(W.copy())
    bias = wk.w(conv.inputs[2]).astype(np.float64) if len(conv.inputs) > 2 and conv.inputs[2] else None
    out = conv.outputs[0]
    nxt = wk.users(out)
    if len(nxt) == 1 and nxt[0].op == "BatchNormalization":
        _take_bn(wk, nxt[0], sd, bn_prefix)
        if bias is not None:   # conv bias in front of a real BN: y = a (x + bias - mean) + b  ->  fold into the mean
            sd[bn_prefix + ".running_mean"] = sd[bn_prefix + ".running_mean"] - torch.from_numpy(bias)
        return nxt[0].outputs[0]
    _identity_bn(sd, bn_prefix, W.shape[0], bias)
    return out


def onnx_to_iresnet_state_dict(path: str):
    """-> (state dict in insightface iresnet naming, architecture 'ir_50' | 'ir_101')."""
    import torch

    from .weights import UNITS
    nodes, inits, inputs = read_onnx(path)
    if len(inputs) != 1:
        raise ValueError(f"{path}: expected one graph input, found {inputs}")
    wk = _Walker(nodes, inits)
    sd: Dict[str, "torch.Tensor"] = {}
    # ---- stem
    stem = wk.only_user(inputs[0], "stem conv")
    t = _conv_then_bn(wk, stem, sd, "conv1.weight", "bn1", 3, 1)
    pre = wk.only_user(t, "stem PReLU")
    if pre.op != "PRelu":
        raise ValueError(f"ONNX import: expected PRelu after the stem, found {pre.op}")
    sd["prelu.weight"] = torch.from_numpy(wk.w(pre.inputs[1]).astype(np.float32).reshape(-1).copy())
    t = pre.outputs[0]
    # ---- residual units until the tail (BN -> Flatten)
    units = []
    while True:
        users = wk.users(t)
        bns = [n for n in users if n.op == "BatchNormalization"]
        if len(users) == 1 and bns:
            after = wk.users(bns[0].outputs[0])
            if len(after) == 1 and after[0].op in ("Flatten", "Reshape"):
                break                                                   # tail reached
        if len(bns) != 1 or len(users) != 2:
            raise ValueError(f"ONNX import: unit {len(units)}: unexpected consumers {[n.op for n in users]} of {t!r}")
        u = len(units)
        other = [n for n in users if n is not bns[0]][0]               # the residual Add itself, or the downsample conv
        # main branch: BN1 -> Conv1 [-> BN2] -> PRelu -> Conv2 [-> BN3] -> Add.  Keys are stage-relative; the stage
        # numbering follows once the number of units (= the architecture) is known.
        tmp: Dict[str, "torch.Tensor"] = {}
        _take_bn(wk, bns[0], tmp, "bn1")
        conv1 = wk.only_user(bns[0].outputs[0], "conv1")
        a = _conv_then_bn(wk, conv1, tmp, "conv1.weight", "bn2", 3, 1)
        pr = wk.only_user(a, "unit PReLU")
        if pr.op != "PRelu":
            raise ValueError(f"ONNX import: unit {u}: expected PRelu, found {pr.op}")
        tmp["prelu.weight"] = torch.from_numpy(wk.w(pr.inputs[1]).astype(np.float32).reshape(-1).copy())
        conv2 = wk.only_user(pr.outputs[0], "conv2")
        stride = int(conv2.attrs.get("strides", [1, 1])[0])
        b = _conv_then_bn(wk, conv2, tmp, "conv2.weight", "bn3", 3, stride)
        add = wk.only_user(b, "residual add")
        if add.op != "Add":
            raise ValueError(f"ONNX import: unit {u}: expected Add, found {add.op}")
        sc_in = [wk.source(i) for i in add.inputs if wk.source(i) != wk.source(b)]
        if len(sc_in) != 1:
            raise ValueError(f"ONNX import: unit {u}: cannot identify the shortcut input of {add.name}")
        if sc_in[0] != t:                                               # downsample: Conv1x1(stride) [-> BN]
            if other.op != "Conv":
                raise ValueError(f"ONNX import: unit {u}: shortcut starts with {other.op}")
            s_out = _conv_then_bn(wk, other, tmp, "downsample.0.weight", "downsample.1", 1, stride)
            if wk.source(s_out) != sc_in[0]:
                raise ValueError(f"ONNX import: unit {u}: shortcut does not end in the residual add")
        elif other is not add:
            raise ValueError(f"ONNX import: unit {u}: identity shortcut expected")
        units.append((tmp, stride))
        t = add.outputs[0]
    arch = {sum(v): k for k, v in UNITS.items()}.get(len(units))
    if arch is None:
        raise ValueError(f"{path}: {len(units)} residual units; iresnet50 has 24, iresnet100 has 49")
    it = iter(units)
    for stage, n in enumerate(UNITS[arch]):
        for j in range(n):
            tmp, stride = next(it)
            if stride != (2 if j == 0 else 1) or (("downsample.0.weight" in tmp) != (j == 0)):
                raise ValueError(f"{path}: unit layer{stage + 1}.{j} has stride {stride} / downsample "
                                 f"{'downsample.0.weight' in tmp}: not the insightface iresnet layout")
            for k, v in tmp.items():
                sd[f"layer{stage + 1}.{j}.{k}"] = v
    # ---- tail: BN2d -> Flatten -> Gemm | MatMul + Add -> [BN1d]
    bn2 = wk.only_user(t, "output BN")
    _take_bn(wk, bn2, sd, "bn2")
    flat = wk.only_user(bn2.outputs[0], "flatten")
    fc = wk.only_user(flat.outputs[0], "fc")
    if fc.op == "Gemm":
        W = wk.w(fc.inputs[1]).astype(np.float64) * float(fc.attrs.get("alpha", 1.0))
        if not int(fc.attrs.get("transB", 0)):
            W = W.T
        bias = wk.w(fc.inputs[2]).astype(np.float64) * float(fc.attrs.get("beta", 1.0)) if len(fc.inputs) > 2 else np.zeros(W.shape[0])
        out = fc.outputs[0]
    elif fc.op == "MatMul":
        W = wk.w(fc.inputs[1]).astype(np.float64).T
        add = wk.only_user(fc.outputs[0], "fc bias")
        if add.op != "Add":
            raise ValueError(f"ONNX import: expected Add after the fc MatMul, found {add.op}")
        bias = wk.w([i for i in add.inputs if i in inits][0]).astype(np.float64)
        out = add.outputs[0]
    else:
        raise ValueError(f"ONNX import: expected Gemm / MatMul for the embedding layer, found {fc.op}")
    if W.shape != (512, 512 * 49):
        raise ValueError(f"{path}: embedding layer is {W.shape}, expected (512, 25088)")
    sd["fc.weight"] = torch.from_numpy(W.astype(np.float32))
    sd["fc.bias"] = torch.from_numpy(bias.astype(np.float32))
    last = wk.users(out)
    if len(last) == 1 and last[0].op == "BatchNormalization":
        _take_bn(wk, last[0], sd, "features")
    elif not last:
        _identity_bn(sd, "features", 512)
    else:
        raise ValueError(f"ONNX import: unexpected nodes after the embedding layer: {[n.op for n in last]}")
    return sd, arch
