"""Warp + normalise and preprocess kernels: byte-exact against cv2 / the reference arithmetic."""

import cv2
import numpy as np
import pytest
import torch

from oracle import align as oa, preprocess as op

pytestmark = pytest.mark.gpu


def _bf16_of_f32(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to(torch.bfloat16)


def _case(rng, S, H=200, W=240):
    src = cv2.GaussianBlur(rng.integers(0, 256, (H, W, 3), dtype=np.uint8), (0, 0), 2)
    tpl = oa.template(S)
    ang = np.deg2rad(rng.uniform(-25, 25))
    sc = rng.uniform(0.6, 2.2) * 112 / S
    R = np.array([[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]]) * sc
    lm = (tpl - S / 2) @ R.T + np.array([W / 2 + rng.uniform(-70, 70), H / 2 + rng.uniform(-70, 70)]) + rng.normal(0, 0.5, (5, 2))
    return src, lm


@pytest.mark.parametrize("S", [112, 224])
def test_align_is_bit_exact_with_reference_align(S):
    """FaceAligner.align (estimateAffinePartial2D + warpAffine) — including crops that leave the frame."""
    from facerecognitionpipeline_b200.face_recognition import FaceAligner
    rng = np.random.default_rng(100 + S)
    al = FaceAligner(output_size=S)
    for _ in range(6):
        src, lm = _case(rng, S)
        ref = oa.align(src, lm, S)
        got = al.align(src, lm)
        assert got.shape == (S, S, 3) and got.dtype == np.uint8
        assert np.array_equal(got, ref)


def test_batch_of_faces_and_fused_normalise():
    from facerecognitionpipeline_b200.face_recognition import FaceAligner
    rng = np.random.default_rng(5)
    al = FaceAligner(output_size=112)
    src = cv2.GaussianBlur(rng.integers(0, 256, (720, 1280, 3), dtype=np.uint8), (0, 0), 2)
    lms = [_case(rng, 112, 720, 1280)[1] for _ in range(9)]
    u8, bf = al.align_batch(src, lms, want_u8=True, want_bf16=True)
    for i, lm in enumerate(lms):
        ref = oa.align(src, lm, 112)
        assert np.array_equal(u8[i], ref)
        exp = _bf16_of_f32(op.preprocess(ref, "adaface")[0].transpose(1, 2, 0))   # NHWC, BGR
        assert torch.equal(bf[i].cpu(), exp)


@pytest.mark.parametrize("S", [112, 224])
def test_preprocess_u8_matches_reference_preprocess(ctx, S):
    rng = np.random.default_rng(S)
    B = 5
    imgs = rng.integers(0, 256, (B, S, S, 3), dtype=np.uint8)
    d_in = torch.from_numpy(imgs).cuda()
    out = torch.empty((2 * B, 112, 112, 3), dtype=torch.bfloat16, device="cuda")
    ctx.frb_preprocess_u8(d_in.data_ptr(), B, S, out.data_ptr(), 1, None)
    torch.cuda.synchronize()
    for b in range(B):
        exp = op.preprocess(imgs[b], "adaface")[0].transpose(1, 2, 0)
        assert torch.equal(out[b].cpu(), _bf16_of_f32(exp))
        flipped = op.preprocess(cv2.flip(imgs[b], 1), "adaface")[0].transpose(1, 2, 0)
        assert torch.equal(out[B + b].cpu(), _bf16_of_f32(flipped))


def test_bad_sizes_are_rejected(ctx):
    from facerecognitionpipeline_b200._native import NativeError
    d = torch.zeros(100, dtype=torch.uint8, device="cuda")
    with pytest.raises(NativeError):
        ctx.frb_preprocess_u8(d.data_ptr(), 1, 160, d.data_ptr(), 0, None)
