"""Drop-in for the hot-path part of the reference `face_recognition` module.

Only `FaceAligner` (face_recognition.py:50-75) is on the hot path: a 5-landmark similarity warp to the
reference's own template followed by cv2.warpAffine(INTER_LINEAR, BORDER_CONSTANT 0).  Here the
resampling runs on the GPU (`frb_warp_normalize`) as a bit-exact restatement of OpenCV's
fixed-point warpAffine, optionally fused with the BGR normalisation into the NHWC bf16 tensor the
backbone consumes.  The 2x3 matrix still comes from cv2.estimateAffinePartial2D on the host
(RANSAC + LM refine on 5 points, microseconds) so that the matrix is the reference's own.

Face detection (insightface buffalo_l, face_recognition.py:19-48) is out of scope: `FaceProcessor`
takes a caller-supplied detector.  `FaceQualityFilter` (face_recognition.py:77-158) is host glue between detector and
embedder that decides WHICH crops reach the hot path (enrollment drops blurred / profile / tiny faces through it), so
it is kept with the reference's arithmetic and ordering, pinned by the reference's own outputs
(tests/golden/make_golden_r2.py).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _native


def similarity_template(output_size: int) -> np.ndarray:
    """The reference's own 5-point template (NOT the canonical ArcFace one), face_recognition.py:53-59."""
    frac = np.array([[0.34, 0.46], [0.66, 0.46], [0.50, 0.61], [0.37, 0.74], [0.63, 0.74]])
    return np.array([[fx * output_size, fy * output_size] for fx, fy in frac], dtype=np.float32)


def estimate_matrix(landmarks: np.ndarray, template: np.ndarray, method: str = "similarity") -> np.ndarray:
    import cv2
    landmarks = np.asarray(landmarks).astype(np.float32)
    if method == "similarity":
        return cv2.estimateAffinePartial2D(landmarks, template)[0]
    return cv2.getAffineTransform(landmarks[:3], template[:3])


class FaceAligner:
    def __init__(self, output_size=112, device: int = 0):
        self.output_size = output_size
        self.template = similarity_template(output_size)
        self._device = device
        self._ctx = None

    def _context(self):
        if self._ctx is None:
            self._ctx = _native.default_context(self._device)
        return self._ctx

    def align_batch(self, image: np.ndarray, landmarks_list: Sequence[np.ndarray], method="similarity",
                    matrices: Optional[Sequence[np.ndarray]] = None, want_u8=True, want_bf16=False):
        """Warp every face of one RGB uint8 frame.  Returns (aligned_u8 [B,S,S,3] or None,
        normalised NHWC bf16 torch tensor [B,112,112,3] on the device or None)."""
        import torch
        image = np.ascontiguousarray(image)
        if image.ndim != 3 or image.shape[2] != 3 or image.dtype != np.uint8:
            raise ValueError("align expects an HxWx3 uint8 RGB image")
        if matrices is None:
            matrices = [estimate_matrix(lm, self.template, method) for lm in landmarks_list]
        B = len(matrices)
        S = self.output_size
        if B == 0:
            return (np.zeros((0, S, S, 3), np.uint8) if want_u8 else None), None
        H, W = image.shape[:2]
        jobs = (_native.WarpJob * B)()
        for i, M in enumerate(matrices):
            if M is None:
                raise ValueError("estimateAffinePartial2D failed for a face")
            jobs[i].src_off, jobs[i].H, jobs[i].W, jobs[i].pitch = 0, H, W, W * 3
            for j, v in enumerate(np.asarray(M, dtype=np.float64).reshape(6)):
                jobs[i].M[j] = float(v)
        ctx = self._context()
        dev = torch.device("cuda", self._device)
        src = torch.from_numpy(image).to(dev)
        out_u8 = torch.empty((B, S, S, 3), dtype=torch.uint8, device=dev) if want_u8 else None
        out_bf = torch.empty((B, 112, 112, 3), dtype=torch.bfloat16, device=dev) if want_bf16 else None
        stream = torch.cuda.current_stream(dev).cuda_stream
        ctx.frb_warp_normalize(src.data_ptr(), jobs, B, S, out_u8.data_ptr() if want_u8 else None,
                               out_bf.data_ptr() if want_bf16 else None, C.c_void_p(stream))
        res_u8 = out_u8.cpu().numpy() if want_u8 else None
        if not want_u8:
            torch.cuda.current_stream(dev).synchronize()
        return res_u8, out_bf

    def align(self, image: np.ndarray, landmarks: np.ndarray, method="similarity") -> np.ndarray:
        return self.align_batch(image, [landmarks], method)[0][0]


class FaceQualityFilter:
    """Same gates, same order, same metrics dict as the reference (face_recognition.py:77-158): detection score, face
    size = min(bbox width, height), yaw / pitch / roll estimated from the 5 landmarks, Laplacian-variance blur score
    of the ALIGNED crop.  The metrics dict only holds what was computed before the first failing gate - callers sort by
    `quality_metrics.get('blur_score', 1000)`, so which keys exist is observable."""

    def __init__(self, min_det_score=0.6, min_face_size=60, max_yaw=45, max_pitch=30, max_roll=30, check_blur=True,
                 blur_threshold=100):
        self.min_det_score = min_det_score
        self.min_face_size = min_face_size
        self.max_yaw, self.max_pitch, self.max_roll = max_yaw, max_pitch, max_roll
        self.check_blur = check_blur
        self.blur_threshold = blur_threshold

    def compute_blur_score(self, face_image: np.ndarray) -> float:
        import cv2
        gray = cv2.cvtColor(face_image, cv2.COLOR_RGB2GRAY) if face_image.ndim == 3 else face_image
        return cv2.Laplacian(gray, cv2.CV_64F).var()

    def compute_pose_angles(self, landmarks: np.ndarray) -> Dict[str, float]:
        # landmark order: left eye, right eye, nose, left mouth corner, right mouth corner; arithmetic stays in the
        # array's own dtype (f32 from the detector) exactly as the reference's expressions evaluate
        eye_l, eye_r, nose, mouth_l, mouth_r = (landmarks[i] for i in range(5))
        eyes_mid = (eye_l + eye_r) / 2
        eyes_vec = eye_r - eye_l
        roll = np.degrees(np.arctan2(eyes_vec[1], eyes_vec[0]))
        dx = nose[0] - eyes_mid[0]
        yaw = np.degrees(np.arcsin(np.clip(dx / np.linalg.norm(eyes_vec), -1, 1))) * 2
        mouth_mid = (mouth_l + mouth_r) / 2
        pitch = ((nose[1] - eyes_mid[1]) / (mouth_mid[1] - eyes_mid[1]) - 0.5) * 60
        return {"yaw": yaw, "pitch": pitch, "roll": roll}

    def is_valid(self, face_dict: Dict, face_image: Optional[np.ndarray] = None) -> Tuple[bool, Dict]:
        metrics: Dict = {"det_score": face_dict["det_score"]}
        if metrics["det_score"] < self.min_det_score:
            return False, metrics
        x1, y1, x2, y2 = (face_dict["bbox"][i] for i in range(4))
        metrics["face_size"] = min(x2 - x1, y2 - y1)
        if metrics["face_size"] < self.min_face_size:
            return False, metrics
        pose = self.compute_pose_angles(face_dict["landmarks"])
        metrics.update(pose)
        for key, limit in (("yaw", self.max_yaw), ("pitch", self.max_pitch), ("roll", self.max_roll)):
            if abs(pose[key]) > limit:
                return False, metrics
        if self.check_blur and face_image is not None:
            metrics["blur_score"] = self.compute_blur_score(face_image)
            if metrics["blur_score"] < self.blur_threshold:
                return False, metrics
        return True, metrics


class FaceProcessor:
    """detect -> align (device warp, bit-exact vs cv2) -> quality filter -> sort, as face_recognition.py:160-216.
    `detector` must provide detect(image_rgb) -> list of dicts with 'bbox', 'landmarks', 'det_score' like the
    reference's FaceDetector (face_recognition.py:30-48); the reference's own detector (insightface) is out of scope."""

    def __init__(self, output_size=224, det_size=(640, 640), det_thresh=0.5, quality_filter_config: Optional[Dict] = None,
                 providers=None, detector=None):
        if detector is None:
            raise ImportError("FaceProcessor needs a face detector (the reference uses insightface buffalo_l, which "
                              "is outside the B200 hot path); pass detector=...")
        self.detector = detector
        self.aligner = FaceAligner(output_size=output_size)
        self.quality_filter = FaceQualityFilter(**(quality_filter_config or {}))

    def process_image(self, image_path: str, return_all: bool = False) -> List[Dict]:
        import cv2
        image = cv2.imread(image_path)
        if image is None:
            raise ValueError(f"Could not load image: {image_path}")
        return self.process_numpy(cv2.cvtColor(image, cv2.COLOR_BGR2RGB), return_all)

    def process_numpy(self, image_rgb: np.ndarray, return_all: bool = False) -> List[Dict]:
        faces = self.detector.detect(image_rgb)
        if not faces:
            return []
        aligned, _ = self.aligner.align_batch(image_rgb, [f["landmarks"] for f in faces])   # all faces of the frame: one launch
        results = []
        for i, f in enumerate(faces):
            ok, metrics = self.quality_filter.is_valid(f, aligned[i])
            if ok or return_all:
                results.append({"aligned_face": aligned[i], "bbox": f["bbox"], "landmarks": f["landmarks"],
                                "det_score": f["det_score"], "quality_metrics": metrics, "is_valid": ok})
        # stable sort, descending det_score x blur; a face rejected before the blur gate has no blur_score and
        # counts as 1000 (face_recognition.py:206-209) - reproduced, it decides enrollment's top-max_faces selection
        results.sort(key=lambda r: r["det_score"] * r["quality_metrics"].get("blur_score", 1000), reverse=True)
        return results if return_all else results[:1]
