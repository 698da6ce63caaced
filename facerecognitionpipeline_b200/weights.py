"""State dict -> BN-folded bf16 layer program for libfrb200 (`frb_backbone_load`).

Replaces `net.build_model(arch)` + `load_state_dict` (reference face_embedder.py:49-56) and the ONNX
iresnet graph (face_embedder.py:64-88): the architecture knowledge (unit list, shortcut kinds,
state-dict key names of the upstream AdaFace `net.py` and insightface `iresnet.py`) lives here; the
C side only executes the resulting generic program.

Folding (all eval-mode BatchNorm, eps = 1e-5, arithmetic in float64 before the final casts):
  * BN after a conv folds into the conv's output channels (scale into weights, shift into bias).
  * The pre-activation BN in front of the first 3x3 conv of a unit cannot fold naively because the
    zero padding is applied AFTER it: its scale folds into the conv's input channels exactly, and its
    shift becomes a bias that depends on which taps fall inside the image -> a 9-case
    (top/mid/bottom x left/mid/right) bias table selected per output pixel in the epilogue.
  * A Conv1x1(stride)+BN shortcut is appended to the second conv's K dimension (same accumulator);
    a MaxPool2d(1, stride) shortcut is a strided residual read in that conv's epilogue.
  * BN2d -> Flatten(NCHW) -> Linear -> BN1d collapse into one [512 x 25088] matrix whose columns are
    permuted to the NHWC flatten order of the device activations.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List

import torch

from . import _native
from ._native import FRB_OP_CONV, FRB_OP_FC, FRB_OP_STEM, LayerDesc

EPS = 1e-5
UNITS = {"ir_50": [3, 4, 14, 3], "ir_101": [3, 13, 30, 3]}
DEPTHS = [64, 128, 256, 512]


def _unit_table(arch: str, layout: str):
    """[(keys, in_c, depth, stride, has_shortcut_conv)] in execution order."""
    if arch not in UNITS:
        raise ValueError(f"Unknown architecture: {arch}. Available: {list(UNITS.keys())}")
    if layout not in ("adaface", "iresnet"):
        raise ValueError(f"Unknown layout: {layout}")
    table = []
    in_c = 64
    i = 0
    for stage, (depth, n) in enumerate(zip(DEPTHS, UNITS[arch])):
        for j in range(n):
            stride = 2 if j == 0 else 1
            cin = in_c if j == 0 else depth
            if layout == "adaface":
                b = f"body.{i}."
                keys = dict(bn1=b + "res_layer.0", conv1=b + "res_layer.1.weight", bn2=b + "res_layer.2",
                            prelu=b + "res_layer.3.weight", conv2=b + "res_layer.4.weight", bn3=b + "res_layer.5",
                            sc_conv=b + "shortcut_layer.0.weight", sc_bn=b + "shortcut_layer.1")
                has_sc = cin != depth
            else:
                b = f"layer{stage + 1}.{j}."
                keys = dict(bn1=b + "bn1", conv1=b + "conv1.weight", bn2=b + "bn2", prelu=b + "prelu.weight",
                            conv2=b + "conv2.weight", bn3=b + "bn3", sc_conv=b + "downsample.0.weight",
                            sc_bn=b + "downsample.1")
                has_sc = j == 0
            table.append((keys, cin, depth, stride, has_sc))
            i += 1
        in_c = depth
    return table


def _head_keys(layout: str):
    if layout == "adaface":
        return dict(stem_conv="input_layer.0.weight", stem_bn="input_layer.1", stem_prelu="input_layer.2.weight",
                    out_bn="output_layer.0", fc_w="output_layer.3.weight", fc_b="output_layer.3.bias",
                    feat_bn="output_layer.4")
    return dict(stem_conv="conv1.weight", stem_bn="bn1", stem_prelu="prelu.weight", out_bn="bn2",
                fc_w="fc.weight", fc_b="fc.bias", feat_bn="features")


def _bn_affine(sd, prefix):
    """BatchNorm (eval) as y = a*x + b, float64."""
    mean = sd[prefix + ".running_mean"].double()
    var = sd[prefix + ".running_var"].double()
    a = 1.0 / torch.sqrt(var + EPS)
    if prefix + ".weight" in sd:
        a = a * sd[prefix + ".weight"].double()
    b = -mean * a
    if prefix + ".bias" in sd:
        b = b + sd[prefix + ".bias"].double()
    return a, b


def _to_bf16_bytes(t: torch.Tensor) -> bytes:
    return t.to(torch.float32).to(torch.bfloat16).contiguous().view(torch.int16).numpy().tobytes()


def _to_f32_bytes(t: torch.Tensor) -> bytes:
    return t.to(torch.float32).contiguous().numpy().tobytes()


class _Blob:
    def __init__(self):
        self.parts: List[bytes] = []
        self.size = 0

    def add(self, b: bytes) -> int:
        pad = (-self.size) % 256
        if pad:
            self.parts.append(b"\0" * pad)
            self.size += pad
        off = self.size
        self.parts.append(b)
        self.size += len(b)
        return off

    def bytes(self) -> bytes:
        return b"".join(self.parts)


@dataclass
class Program:
    arch: str
    layout: str
    layers: List[LayerDesc]
    blob: bytes
    n_bufs: int
    l2_in_model: bool
    # float32 copies of exactly what was packed (bf16-rounded weights), for tests / emulation
    debug: List[dict]

    def load_into(self, ctx: "_native.Context"):
        arr = (LayerDesc * len(self.layers))(*self.layers)
        buf = C.create_string_buffer(self.blob, len(self.blob))
        ctx.frb_backbone_load(arr, len(self.layers), C.cast(buf, C.c_void_p), len(self.blob), self.n_bufs)


def border_bias_table(T: torch.Tensor, base: torch.Tensor) -> torch.Tensor:
    """T: [9][Cout] per-tap contribution of the folded pre-conv shift; base: [Cout].
    Returns [9][Cout]: case = rowcase*3 + colcase with rowcase/colcase 0 = first, 1 = interior,
    2 = last row/column; a tap (r, s) is dropped when it reads the zero padding."""
    out = torch.zeros(9, T.shape[1], dtype=torch.float64)
    for rc in range(3):
        for cc in range(3):
            acc = base.clone()
            for r in range(3):
                if (rc == 0 and r == 0) or (rc == 2 and r == 2):
                    continue
                for s in range(3):
                    if (cc == 0 and s == 0) or (cc == 2 and s == 2):
                        continue
                    acc = acc + T[r * 3 + s]
            out[rc * 3 + cc] = acc
    return out


def build_program(state_dict: Dict[str, torch.Tensor], arch: str = "ir_101", layout: str = "adaface",
                  keep_debug: bool = False) -> Program:
    sd = state_dict
    unit_table = _unit_table(arch, layout)   # validates arch / layout first
    hk = _head_keys(layout)
    blob = _Blob()
    layers: List[LayerDesc] = []
    debug: List[dict] = []

    def new_layer(**kw) -> LayerDesc:
        L = LayerDesc()
        L.sc_buf = -1
        L.res_buf = -1
        L.bias_cases = 1
        for k, v in kw.items():
            setattr(L, k, v)
        return L

    # ---- stem: Conv3x3(3->64) + BN + PReLU -> [64][32] bf16 K-major, k = (r*3+s)*3 + c, k = 27..31 zero
    a, b = _bn_affine(sd, hk["stem_bn"])
    w = sd[hk["stem_conv"]].double() * a[:, None, None, None]          # [64,3,3,3] (co,ci,r,s)
    w_k = torch.zeros(64, 32, dtype=torch.float64)
    w_k[:, :27] = w.permute(0, 2, 3, 1).reshape(64, 27)                  # co x (r,s,ci)
    L = new_layer(op=FRB_OP_STEM, cin=3, cout=64, hin=112, win=112, ksize=3, stride=1, pad=1, in_buf=-1, out_buf=0,
                  has_prelu=1)
    wb = _to_bf16_bytes(w_k)
    L.w_off, L.w_bytes = blob.add(wb), len(wb)
    L.bias_off = blob.add(_to_f32_bytes(b))
    L.prelu_off = blob.add(_to_f32_bytes(sd[hk["stem_prelu"]].double()))
    layers.append(L)
    if keep_debug:
        w_t = w_k[:, :27].t().float().to(torch.bfloat16).float()        # [27][64], the values the device multiplies
        debug.append(dict(kind="stem", w=w_t, bias=b.float(), prelu=sd[hk["stem_prelu"]].float()))

    X, Hb, Y = 0, 1, 2  # activation buffers: unit input, conv1 output, unit output
    h = 112
    for keys, cin, d, stride, has_sc in unit_table:
        a1, b1 = _bn_affine(sd, keys["bn1"])
        a2, b2 = _bn_affine(sd, keys["bn2"])
        a3, b3 = _bn_affine(sd, keys["bn3"])
        W1 = sd[keys["conv1"]].double()                                   # [d,cin,3,3]
        W1f = W1 * a2[:, None, None, None] * a1[None, :, None, None]
        # per-tap bias of the folded shift: T[tap][co] = a2[co] * sum_ci W1[co,ci,r,s] * b1[ci]
        T = (torch.einsum("oirs,i->rso", W1, b1) * a2[None, None, :]).reshape(9, d)
        table = border_bias_table(T, b2)
        w1_k = W1f.permute(0, 2, 3, 1).reshape(d, 9 * cin)               # K order (r,s,ci)
        L1 = new_layer(op=FRB_OP_CONV, cin=cin, cout=d, hin=h, win=h, ksize=3, stride=1, pad=1, in_buf=X, out_buf=Hb,
                       bias_cases=9, has_prelu=1)
        wb = _to_bf16_bytes(w1_k)
        L1.w_off, L1.w_bytes = blob.add(wb), len(wb)
        L1.bias_off = blob.add(_to_f32_bytes(table))
        L1.prelu_off = blob.add(_to_f32_bytes(sd[keys["prelu"]].double()))
        layers.append(L1)

        W2f = sd[keys["conv2"]].double() * a3[:, None, None, None]
        w2_k = W2f.permute(0, 2, 3, 1).reshape(d, 9 * d)
        bias2 = b3.clone()
        ho = (h + 2 - 3) // stride + 1
        L2 = new_layer(op=FRB_OP_CONV, cin=d, cout=d, hin=h, win=h, ksize=3, stride=stride, pad=1, in_buf=Hb, out_buf=Y,
                       bias_cases=1, has_prelu=0)
        if has_sc:
            asc, bsc = _bn_affine(sd, keys["sc_bn"])
            Wsc = sd[keys["sc_conv"]].double()[:, :, 0, 0] * asc[:, None]  # [d,cin]
            w2_k = torch.cat([w2_k, Wsc], dim=1)
            bias2 = bias2 + bsc
            L2.sc_buf, L2.sc_cin, L2.sc_hin, L2.sc_win, L2.sc_stride = X, cin, h, h, stride
        else:
            L2.res_buf, L2.res_h, L2.res_w, L2.res_stride = X, h, h, stride
        wb = _to_bf16_bytes(w2_k)
        L2.w_off, L2.w_bytes = blob.add(wb), len(wb)
        L2.bias_off = blob.add(_to_f32_bytes(bias2))
        L2.prelu_off = L2.bias_off
        layers.append(L2)
        if keep_debug:
            debug.append(dict(kind="unit", cin=cin, d=d, stride=stride, has_sc=has_sc, h=h,
                              w1=w1_k.to(torch.bfloat16).float(), table=table.float(), prelu=sd[keys["prelu"]].float(),
                              w2=w2_k.to(torch.bfloat16).float(), bias2=bias2.float()))
        X, Y = Y, X
        h = ho

    # ---- tail: BN2d -> Flatten(NCHW) -> Linear -> BN1d  ==> one [512 x 25088] GEMM + bias
    ao, bo = _bn_affine(sd, hk["out_bn"])                                  # [512] over channels
    af, bf = _bn_affine(sd, hk["feat_bn"])                                 # [512] over features
    Wfc = sd[hk["fc_w"]].double().reshape(512, 512, 49)                    # [o][c][hw]
    bias_fc = sd[hk["fc_b"]].double() + torch.einsum("och,c->o", Wfc, bo)
    Wf = Wfc * ao[None, :, None] * af[:, None, None]
    bias_f = bias_fc * af + bf
    w_k = Wf.permute(0, 2, 1).reshape(512, 49 * 512)                       # NHWC flatten: (hw)*512 + c
    Lf = new_layer(op=FRB_OP_FC, cin=49 * 512, cout=512, hin=1, win=1, ksize=1, stride=1, pad=0, in_buf=X, out_buf=Hb)
    wb = _to_bf16_bytes(w_k)
    Lf.w_off, Lf.w_bytes = blob.add(wb), len(wb)
    Lf.bias_off = blob.add(_to_f32_bytes(bias_f))
    Lf.prelu_off = Lf.bias_off
    layers.append(Lf)
    if keep_debug:
        debug.append(dict(kind="fc", w=w_k.to(torch.bfloat16).float(), bias=bias_f.float()))
    return Program(arch=arch, layout=layout, layers=layers, blob=blob.bytes(), n_bufs=3,
                   l2_in_model=(layout == "adaface"), debug=debug)


def load_checkpoint_state_dict(model_path: str) -> Dict[str, torch.Tensor]:
    """AdaFace Lightning checkpoint: ['state_dict'] with the 'model.' prefix stripped
    (reference face_embedder.py:51-53)."""
    ckpt = torch.load(model_path, map_location="cpu", weights_only=False)
    statedict = ckpt["state_dict"] if "state_dict" in ckpt else ckpt
    if any(k.startswith("model.") for k in statedict):
        statedict = {k[6:]: v for k, v in statedict.items() if k.startswith("model.")}
    return statedict


def random_init_state_dict(arch: str, layout: str, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Random-init weights for when no checkpoint is shipped (kaiming-normal fan_out convs/linear as
    upstream `initialize_weights`, randomised BN affine/stats, PReLU slopes U[0.1,0.4]).
    NOTE: unlike the oracle's generator this does not calibrate BN statistics (that needs a CPU
    forward pass); tests and the bench pass the oracle-generated state dict to both sides instead."""
    import math
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    hk = _head_keys(layout)

    def conv(name, co, ci, k):
        sd[name] = torch.randn(co, ci, k, k, generator=g) * math.sqrt(2.0 / (co * k * k))

    def bn(prefix, c, affine=True, gain=1.0):
        if affine:
            sd[prefix + ".weight"] = (torch.rand(c, generator=g) + 0.5) * gain
            sd[prefix + ".bias"] = torch.randn(c, generator=g) * 0.1
        sd[prefix + ".running_mean"] = torch.randn(c, generator=g) * 0.1
        sd[prefix + ".running_var"] = torch.rand(c, generator=g) + 0.5

    conv(hk["stem_conv"], 64, 3, 3)
    bn(hk["stem_bn"], 64)
    sd[hk["stem_prelu"]] = torch.rand(64, generator=g) * 0.3 + 0.1
    for keys, cin, d, stride, has_sc in _unit_table(arch, layout):
        bn(keys["bn1"], cin)
        conv(keys["conv1"], d, cin, 3)
        bn(keys["bn2"], d)
        sd[keys["prelu"]] = torch.rand(d, generator=g) * 0.3 + 0.1
        conv(keys["conv2"], d, d, 3)
        bn(keys["bn3"], d, gain=0.25)  # keep the residual stream from growing geometrically
        if has_sc:
            conv(keys["sc_conv"], d, cin, 1)
            bn(keys["sc_bn"], d)
    bn(hk["out_bn"], 512)
    sd[hk["fc_w"]] = torch.randn(512, 512 * 49, generator=g) * math.sqrt(2.0 / 512)
    sd[hk["fc_b"]] = torch.randn(512, generator=g) * 0.1
    bn(hk["feat_bn"], 512, affine=(layout == "iresnet"))
    return sd
