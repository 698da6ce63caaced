import os

import cv2
import numpy as np

from oracle import align as oa, preprocess as op


def test_preprocess_matches_reference_arithmetic(golden_dir):
    z = np.load(os.path.join(golden_dir, "preprocess_cases.npz"))
    for i in range(int(z["num_cases"])):
        img = z[f"p{i}/img"]
        np.testing.assert_array_equal(op.preprocess(img, "adaface"), z[f"p{i}/adaface"])
        np.testing.assert_array_equal(op.preprocess(img, "arcface"), z[f"p{i}/arcface"])


def test_normalisation_is_a_256_entry_lut():
    """(x/255-0.5)/0.5 and (x-127.5)/127.5 give bit-identical float32 for all byte values,
    which is what lets the kernel use one 256-entry table."""
    x = np.arange(256, dtype=np.uint8)
    a = ((x / 255.0 - 0.5) / 0.5).astype(np.float32)
    b = ((x - 127.5) / 127.5).astype(np.float32)
    assert np.array_equal(a, b)


def test_resize_224_to_112_is_box_mean_round_half_up():
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (224, 224, 3), dtype=np.uint8)
    ref = cv2.resize(img, (112, 112), interpolation=cv2.INTER_LINEAR)
    t = img.astype(np.int32)
    box = (t[0::2, 0::2] + t[0::2, 1::2] + t[1::2, 0::2] + t[1::2, 1::2] + 2) >> 2
    assert np.array_equal(ref, box.astype(np.uint8))


def _case(rng, S):
    src = cv2.GaussianBlur(rng.integers(0, 256, (200, 240, 3), dtype=np.uint8), (0, 0), 2)
    tpl = oa.template(S)
    ang = np.deg2rad(rng.uniform(-25, 25))
    sc = rng.uniform(0.6, 2.2) * 112 / S
    R = np.array([[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]]) * sc
    lm = (tpl - S / 2) @ R.T + np.array([120 + rng.uniform(-70, 70), 100 + rng.uniform(-70, 70)]) + rng.normal(0, 0.5, (5, 2))
    return src, lm


def test_template_values():
    t = oa.template(112)
    np.testing.assert_allclose(t, [[38.08, 51.52], [73.92, 51.52], [56, 68.32], [41.44, 82.88], [70.56, 82.88]], atol=1e-4)


def test_fixed_point_warp_is_bit_exact_with_cv2():
    rng = np.random.default_rng(11)
    for S in (112, 112, 112, 224):
        src, lm = _case(rng, S)
        M = oa.estimate(lm, S)
        ref = oa.align(src, lm, S)
        got = oa.warp_affine_fixed_point(src, M, S)
        assert np.array_equal(ref, got)
