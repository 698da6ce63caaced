"""CPU oracle of FaceAligner.align — TEST INFRASTRUCTURE ONLY (see oracle/backbone.py header).

`align` executes the reference's own calls (face_recognition.py:53-75): template = fractions * S,
cv2.estimateAffinePartial2D(landmarks, template)[0], cv2.warpAffine(INTER_LINEAR, BORDER_CONSTANT, 0).
`warp_affine_fixed_point` is a plain-numpy restatement of what cv::warpAffine computes for 8-bit
INTER_LINEAR (imgwarp.cpp: AB_BITS = 10, INTER_BITS = 5, 15-bit bilinear weight table whose four
weights are forced to sum to 32768, (sum + 2^14) >> 15) — it is checked bit-for-bit against
cv2.warpAffine in tests/test_oracle_align.py and is what the CUDA kernel restates.
"""
import cv2
import numpy as np

TEMPLATE_FRACTIONS = np.array([[0.34, 0.46], [0.66, 0.46], [0.50, 0.61], [0.37, 0.74], [0.63, 0.74]])


def template(output_size: int) -> np.ndarray:
    return np.array([[fx * output_size, fy * output_size] for fx, fy in TEMPLATE_FRACTIONS], dtype=np.float32)


def estimate(landmarks: np.ndarray, output_size: int = 112) -> np.ndarray:
    return cv2.estimateAffinePartial2D(landmarks.astype(np.float32), template(output_size))[0]


def align(image: np.ndarray, landmarks: np.ndarray, output_size: int = 112) -> np.ndarray:
    tform = estimate(landmarks, output_size)
    return cv2.warpAffine(image, tform, (output_size, output_size), flags=cv2.INTER_LINEAR,
                          borderMode=cv2.BORDER_CONSTANT, borderValue=0)


def bilinear_tab() -> np.ndarray:
    tab = np.zeros((32, 32, 4), np.int32)
    for ay in range(32):
        for ax in range(32):
            fx, fy = np.float32(ax / 32.0), np.float32(ay / 32.0)
            one = np.float32(1.0)
            w = np.array([(one - fy) * (one - fx), (one - fy) * fx, fy * (one - fx), fy * fx], np.float32)
            iw = np.clip(np.rint(w * np.float32(32768)), -32768, 32767).astype(np.int32)
            diff = int(iw.sum()) - 32768
            if diff < 0:
                iw[int(np.argmax(iw))] -= diff
            elif diff > 0:
                iw[int(np.argmin(iw))] -= diff
            tab[ay, ax] = iw
    return tab


def warp_affine_fixed_point(src: np.ndarray, M: np.ndarray, S: int) -> np.ndarray:
    M = np.asarray(M, np.float64)
    D = M[0, 0] * M[1, 1] - M[0, 1] * M[1, 0]
    D = 1.0 / D if D != 0 else 0.0
    m00, m11 = M[1, 1] * D, M[0, 0] * D
    m01, m10 = M[0, 1] * (-D), M[1, 0] * (-D)
    b1 = -m00 * M[0, 2] - m01 * M[1, 2]
    b2 = -m10 * M[0, 2] - m11 * M[1, 2]
    sat = lambda v: np.clip(np.rint(v), -2 ** 31, 2 ** 31 - 1).astype(np.int64)
    xs = np.arange(S, dtype=np.float64)
    adelta, bdelta = sat(m00 * xs * 1024), sat(m10 * xs * 1024)
    H, W = src.shape[:2]
    tab = bilinear_tab()
    out = np.zeros((S, S, 3), np.uint8)
    srcp = np.zeros((H + 2, W + 2, 3), np.int64)   # zero border = BORDER_CONSTANT 0
    srcp[1:-1, 1:-1] = src
    for y in range(S):
        X = (int(sat((m01 * y + b1) * 1024)) + 16 + adelta) >> 5
        Y = (int(sat((m11 * y + b2) * 1024)) + 16 + bdelta) >> 5
        sx, sy, ax, ay = X >> 5, Y >> 5, X & 31, Y & 31
        w = tab[ay, ax].astype(np.int64)          # [S,4]
        acc = np.zeros((S, 3), np.int64)
        for k, (dy, dx) in enumerate([(0, 0), (0, 1), (1, 0), (1, 1)]):
            yy, xx = sy + dy, sx + dx
            ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
            px = srcp[np.clip(yy, -1, H) + 1, np.clip(xx, -1, W) + 1]
            acc += np.where(ok[:, None], px, 0) * w[:, k:k + 1]
        out[y] = np.clip((acc + 16384) >> 15, 0, 255)
    return out
