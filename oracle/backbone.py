"""CPU fp32 oracle of the IR-50 / IR-101 face backbone — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product path (facerecognitionpipeline_b200/) never does.

PARITY STATUS: "parity unpinned" for the backbone.  The reference does not contain the backbone
arithmetic: face_embedder.py:11 does `import net` and face_embedder.py:49 calls
`net.build_model(architecture)`, where net.py is the un-vendored, un-pinned upstream repository
mk-minchul/AdaFace; the ArcFace variant is an ONNX export of insightface
`recognition/arcface_torch/backbones/iresnet.py` run by onnxruntime (face_embedder.py:65,78).
Neither file, nor any checkpoint, nor any input/output pair exists under /root/reference, and the
reference has no tests.  This module therefore restates the *published* architectures:

  adaface : Backbone(input_size=(112,112), num_layers in {50,100}, mode='ir')
            input_layer  = Conv3x3(3,64,s1,p1,no bias) -> BN2d -> PReLU(64)
            body[i]      = BasicBlockIR(in, depth, stride):
                             res      = BN2d(in) -> Conv3x3(in,depth,s1,p1) -> BN2d -> PReLU(depth)
                                        -> Conv3x3(depth,depth,stride,p1) -> BN2d
                             shortcut = MaxPool2d(1,stride) if in == depth else Conv1x1(in,depth,stride)+BN2d
                             out      = res + shortcut
            blocks: 50 -> [3,4,14,3], 100 -> [3,13,30,3]; first unit of every stage has stride 2
            output_layer = BN2d(512) -> Dropout(eval: identity) -> Flatten(NCHW) -> Linear(25088,512)
                           -> BN1d(512, affine=False);  forward returns (x/||x||_2, ||x||_2)
  iresnet : conv1/bn1/prelu stem, IBasicBlock with the same op order, `downsample` = Conv1x1+BN on the
            first unit of EVERY stage (incl. 64->64 s2), bn2 -> flatten -> fc -> features(BN1d affine);
            returns the raw feature (no L2 norm inside the model).

Round 2: the arithmetic of both graphs AS RESTATED HERE is checked against an independent engine - the networks
are written out as ONNX (tests/onnx_writer.py) and run by OpenCV's dnn module; this module agrees to ~2e-5 relative
(tests/test_oracle_backbone.py, tests/test_onnx_import.py).  For the ArcFace path that is the reference's own
definition of the arithmetic (an ONNX runtime executing the file).  Whether the AdaFace graph restated here IS
upstream's net.py remains unpinned.

State-dict key names follow upstream so real `adaface_ir*.ckpt` state dicts load unchanged
(face_embedder.py:51-53 strips the `model.` prefix).  All BatchNorms use eps = 1e-5, eval mode.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import numpy as np
import torch
import torch.nn.functional as F

EPS = 1e-5
UNITS = {"ir_50": [3, 4, 14, 3], "ir_101": [3, 13, 30, 3]}
DEPTHS = [64, 128, 256, 512]


def unit_specs(arch: str) -> List[Tuple[int, int, int]]:
    """[(in_channel, depth, stride)] for every residual unit, upstream get_blocks()."""
    if arch not in UNITS:
        raise ValueError(f"Unknown architecture: {arch}. Available: {list(UNITS.keys())}")
    specs = []
    in_c = 64
    for depth, n in zip(DEPTHS, UNITS[arch]):
        specs.append((in_c, depth, 2))
        specs += [(depth, depth, 1)] * (n - 1)
        in_c = depth
    return specs


# ----------------------------------------------------------------------------- key naming
def adaface_keys(i: int):
    b = f"body.{i}."
    return dict(bn1=b + "res_layer.0", conv1=b + "res_layer.1.weight", bn2=b + "res_layer.2",
                prelu=b + "res_layer.3.weight", conv2=b + "res_layer.4.weight", bn3=b + "res_layer.5",
                sc_conv=b + "shortcut_layer.0.weight", sc_bn=b + "shortcut_layer.1")


def iresnet_keys(stage: int, j: int):
    b = f"layer{stage + 1}.{j}."
    return dict(bn1=b + "bn1", conv1=b + "conv1.weight", bn2=b + "bn2", prelu=b + "prelu.weight",
                conv2=b + "conv2.weight", bn3=b + "bn3", sc_conv=b + "downsample.0.weight",
                sc_bn=b + "downsample.1")


def unit_key_list(arch: str, layout: str):
    """Per-unit key dict + (in, depth, stride, has_sc_conv)."""
    out = []
    specs = unit_specs(arch)
    if layout == "adaface":
        for i, (in_c, d, s) in enumerate(specs):
            out.append((adaface_keys(i), in_c, d, s, in_c != d))
    elif layout == "iresnet":
        i = 0
        for stage, n in enumerate(UNITS[arch]):
            for j in range(n):
                in_c, d, s = specs[i]
                out.append((iresnet_keys(stage, j), in_c, d, s, j == 0))
                i += 1
    else:
        raise ValueError(f"Unknown layout: {layout}")
    return out


def head_keys(layout: str):
    if layout == "adaface":
        return dict(stem_conv="input_layer.0.weight", stem_bn="input_layer.1", stem_prelu="input_layer.2.weight",
                    out_bn="output_layer.0", fc_w="output_layer.3.weight", fc_b="output_layer.3.bias",
                    feat_bn="output_layer.4")
    return dict(stem_conv="conv1.weight", stem_bn="bn1", stem_prelu="prelu.weight", out_bn="bn2",
                fc_w="fc.weight", fc_b="fc.bias", feat_bn="features")


# ----------------------------------------------------------------------------- weights
def random_state_dict(arch: str = "ir_101", layout: str = "adaface", seed: int = 0,
                      calibrate: bool = True) -> Dict[str, torch.Tensor]:
    """Random-init weights of the named architecture (no checkpoints are shipped):
    upstream initialiser for conv / linear (kaiming normal, fan_out) PLUS randomised BatchNorm
    running stats / affine and PReLU slopes, so that BN-folding mistakes cannot hide behind
    identity BatchNorms."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def conv(name, co, ci, k):
        std = math.sqrt(2.0 / (co * k * k))
        sd[name] = torch.randn(co, ci, k, k, generator=g) * std

    def bn(prefix, c, affine=True):
        if affine:
            sd[prefix + ".weight"] = torch.rand(c, generator=g) + 0.5
            sd[prefix + ".bias"] = torch.randn(c, generator=g) * 0.1
        sd[prefix + ".running_mean"] = torch.randn(c, generator=g) * 0.1
        sd[prefix + ".running_var"] = torch.rand(c, generator=g) + 0.5
        sd[prefix + ".num_batches_tracked"] = torch.tensor(0)

    def prelu(name, c):
        sd[name] = torch.rand(c, generator=g) * 0.3 + 0.1

    hk = head_keys(layout)
    conv(hk["stem_conv"], 64, 3, 3)
    bn(hk["stem_bn"], 64)
    prelu(hk["stem_prelu"], 64)
    for keys, in_c, d, s, has_sc in unit_key_list(arch, layout):
        bn(keys["bn1"], in_c)
        conv(keys["conv1"], d, in_c, 3)
        bn(keys["bn2"], d)
        prelu(keys["prelu"], d)
        conv(keys["conv2"], d, d, 3)
        bn(keys["bn3"], d)
        if has_sc:
            conv(keys["sc_conv"], d, in_c, 1)
            bn(keys["sc_bn"], d)
    bn(hk["out_bn"], 512)
    sd[hk["fc_w"]] = torch.randn(512, 512 * 49, generator=g) * math.sqrt(2.0 / 512)
    sd[hk["fc_b"]] = torch.randn(512, generator=g) * 0.1
    bn(hk["feat_bn"], 512, affine=(layout == "iresnet"))
    if calibrate:
        calibrate_bn_(sd, arch, layout, seed)
    return sd


# ----------------------------------------------------------------------------- forward
def _bn(sd, prefix, x):
    w = sd.get(prefix + ".weight")
    b = sd.get(prefix + ".bias")
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"], w, b, False, 0.0, EPS)


def _run(sd, x, arch, layout, bn, return_intermediates=False):
    hk = head_keys(layout)
    inter = []
    x = F.conv2d(x, sd[hk["stem_conv"]], None, 1, 1)
    x = F.prelu(bn(sd, hk["stem_bn"], x), sd[hk["stem_prelu"]])
    if return_intermediates:
        inter.append(x)
    for keys, in_c, d, s, has_sc in unit_key_list(arch, layout):
        if has_sc:
            shortcut = bn(sd, keys["sc_bn"], F.conv2d(x, sd[keys["sc_conv"]], None, s, 0))
        else:
            shortcut = F.max_pool2d(x, 1, s)
        r = bn(sd, keys["bn1"], x)
        r = F.conv2d(r, sd[keys["conv1"]], None, 1, 1)
        r = F.prelu(bn(sd, keys["bn2"], r), sd[keys["prelu"]])
        r = F.conv2d(r, sd[keys["conv2"]], None, s, 1)
        r = bn(sd, keys["bn3"], r)
        x = r + shortcut
        if return_intermediates:
            inter.append(x)
    x = bn(sd, hk["out_bn"], x)
    x = x.reshape(x.shape[0], -1)  # NCHW flatten: index c*49 + h*7 + w
    x = F.linear(x, sd[hk["fc_w"]], sd[hk["fc_b"]])
    x = bn(sd, hk["feat_bn"], x)
    if layout == "adaface":
        norm = torch.norm(x, 2, 1, True)
        out = (torch.div(x, norm), norm)
    else:
        out = x
    return (out, inter) if return_intermediates else out


@torch.no_grad()
def forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, arch: str = "ir_101", layout: str = "adaface",
            return_intermediates: bool = False):
    """x: [B,3,112,112] fp32 (BGR, normalised as FaceEmbedder.preprocess does).
    adaface -> (features/||features||, ||features||); iresnet -> raw features [B,512]."""
    return _run(sd, x, arch, layout, _bn, return_intermediates)


@torch.no_grad()
def calibrate_bn_(sd, arch, layout, seed=0, batch=8):
    """Set every BatchNorm's running statistics to the statistics of a synthetic calibration batch
    (jittered), so that activations stay O(1) through the 24/49 units the way they do in a trained
    checkpoint; otherwise a random-init residual net grows geometrically and biases stop mattering."""
    g = torch.Generator().manual_seed(seed + 7919)
    x = torch.randint(0, 256, (batch, 3, 112, 112), generator=g).float()
    x = F.avg_pool2d(x, 3, 1, 1)
    x = (x / 255.0 - 0.5) / 0.5

    def bn_cal(sd_, prefix, t):
        dims = [0] + list(range(2, t.dim()))
        mean = t.mean(dims)
        var = t.var(dims, unbiased=False)
        c = mean.numel()
        sd_[prefix + ".running_mean"] = mean + torch.randn(c, generator=g) * 0.1 * var.sqrt()
        sd_[prefix + ".running_var"] = var * (torch.rand(c, generator=g) * 0.5 + 0.75) + 1e-3
        return _bn(sd_, prefix, t)

    _run(sd, x, arch, layout, bn_cal)
    return sd


class OracleModel:
    """Callable with the seam the reference uses: model(x) -> (features, norm) (face_embedder.py:119)."""

    def __init__(self, arch="ir_101", layout="adaface", state_dict=None, seed=0):
        self.arch, self.layout = arch, layout
        self.sd = state_dict if state_dict is not None else random_state_dict(arch, layout, seed)

    def __call__(self, x):
        return forward(self.sd, x, self.arch, self.layout)


def flops_per_face(arch: str, layout: str = "adaface") -> float:
    """2*MACs of convs + FC (BN/PReLU excluded), SURVEY §8a."""
    macs = 112 * 112 * 64 * 27
    h = 112
    for keys, in_c, d, s, has_sc in unit_key_list(arch, layout):
        macs += h * h * d * in_c * 9
        ho = h // s
        macs += ho * ho * d * d * 9
        if has_sc:
            macs += ho * ho * d * in_c
        h = ho
    macs += 25088 * 512
    return 2.0 * macs
