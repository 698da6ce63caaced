"""CPU emulation of the *packed device program* (facerecognitionpipeline_b200.weights.Program) —
TEST INFRASTRUCTURE ONLY.  It executes the BN-folded layers with fp32 accumulation and rounds
activations to bf16 exactly where the kernels store them, so a test can separate "the folding is
wrong" (emulation disagrees with oracle.backbone) from "a kernel is wrong" (device disagrees with
the emulation)."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _bf16(x):
    return x.to(torch.bfloat16).to(torch.float32)


def _case_map(P, Q):
    rc = torch.ones(P, dtype=torch.long)
    rc[0], rc[-1] = 0, 2
    cc = torch.ones(Q, dtype=torch.long)
    cc[0], cc[-1] = 0, 2
    return rc[:, None] * 3 + cc[None, :]


@torch.no_grad()
def run_program(program, x_nchw_f32: torch.Tensor, quantize: bool = True, return_intermediates: bool = False):
    """x: [B,3,112,112] fp32 preprocessed (BGR).  Returns raw FC output [B,512] (before L2)."""
    q = _bf16 if quantize else (lambda t: t)
    dbg = program.debug
    assert dbg, "build_program(..., keep_debug=True) required"
    inter = []
    x = q(x_nchw_f32)
    st = dbg[0]
    w = st["w"].reshape(3, 3, 3, 64).permute(3, 2, 0, 1)  # (r,s,ci,co) -> (co,ci,r,s)
    y = F.conv2d(x, w, st["bias"], 1, 1)
    y = torch.where(y > 0, y, y * st["prelu"][None, :, None, None])
    x = q(y)
    if return_intermediates:
        inter.append(x)
    for u in dbg[1:-1]:
        cin, d, s, h = u["cin"], u["d"], u["stride"], u["h"]
        w1 = u["w1"].reshape(d, 3, 3, cin).permute(0, 3, 1, 2)
        y = F.conv2d(x, w1, None, 1, 1)
        table = u["table"][_case_map(h, h)]            # [h,h,d]
        y = y + table.permute(2, 0, 1)[None]
        y = torch.where(y > 0, y, y * u["prelu"][None, :, None, None])
        hmid = q(y)
        w2 = u["w2"][:, :9 * d].reshape(d, 3, 3, d).permute(0, 3, 1, 2)
        y = F.conv2d(hmid, w2, None, s, 1)
        if u["has_sc"]:
            wsc = u["w2"][:, 9 * d:].reshape(d, cin, 1, 1)
            y = y + F.conv2d(x, wsc, None, s, 0)
        y = y + u["bias2"][None, :, None, None]
        if not u["has_sc"]:
            y = y + x[:, :, ::s, ::s]
        x = q(y)
        if return_intermediates:
            inter.append(x)
    fc = dbg[-1]
    flat = x.permute(0, 2, 3, 1).reshape(x.shape[0], -1)  # NHWC flatten
    out = flat @ fc["w"].t() + fc["bias"]
    return (out, inter) if return_intermediates else out
