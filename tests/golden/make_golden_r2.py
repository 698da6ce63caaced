#!/usr/bin/env python
"""Golden fixtures (round 2) produced by the REFERENCE's own code, run in the build container only (the GPU box never
sees /root/reference).  Classes / methods are pulled out of the source files with `ast` and executed unmodified
(the modules themselves do not import: insightface / flask are missing).

  r2_cases.npz
    quality/*     FaceQualityFilter.is_valid / compute_pose_angles / compute_blur_score (face_recognition.py:77-158)
                  on seeded synthetic detections + aligned crops of varying sharpness
    process/*     FaceProcessor.process_numpy (face_recognition.py:184-216) with a synthetic detector and the
                  reference's own FaceAligner: which detections come back, in which order, with which is_valid flag
    bestframe/*   LiveRecognitionTracker.get_best_frame and the should_recognize gate
                  (face_recognition_server.py:39-85) on seeded per-track frame buffers
"""
import ast
import os
import types
from collections import deque

import cv2
import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def extract_class(path, cls, drop=()):
    tree = ast.parse(open(path).read())
    node = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls)
    node.body = [n for n in node.body if not (isinstance(n, ast.FunctionDef) and n.name in drop)]
    from typing import Dict, List, Optional, Tuple
    import time
    from datetime import datetime
    ns = dict(cv2=cv2, np=np, List=List, Dict=Dict, Tuple=Tuple, Optional=Optional, deque=deque, time=time, datetime=datetime)
    exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    return ns[cls]


def synth_faces(rng, n, W=640, H=480):
    """Detections the way insightface reports them: bbox f32 [x1,y1,x2,y2], 5 landmarks f32, det_score f32."""
    tpl = np.array([[0.34, 0.46], [0.66, 0.46], [0.50, 0.61], [0.37, 0.74], [0.63, 0.74]])
    faces = []
    for _ in range(n):
        size = float(rng.choice([40, 58, 60, 61, 90, 140, 200]))
        cx, cy = rng.uniform(120, W - 120), rng.uniform(120, H - 120)
        ang = np.deg2rad(float(rng.choice([0, 5, -12, 29, 31, -35, 50])))
        R = np.array([[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]])
        lm = (tpl - 0.5) * size
        lm[2, 0] += size * float(rng.choice([0, 0.02, 0.08, 0.12, -0.15]))      # nose sideways: yaw
        lm[2, 1] += size * float(rng.choice([0, 0.03, -0.06, 0.09]))            # nose up/down: pitch
        lm = lm @ R.T + np.array([cx, cy]) + rng.normal(0, 0.3, (5, 2))
        aspect = float(rng.choice([1.0, 0.8, 1.3]))
        bbox = np.array([cx - size / 2, cy - size * aspect / 2, cx + size / 2, cy + size * aspect / 2], np.float32)
        faces.append(dict(bbox=bbox, landmarks=lm.astype(np.float32),
                          det_score=np.float32(rng.choice([0.45, 0.59, 0.6, 0.61, 0.75, 0.9, 0.97]))))
    return faces


def main():
    pack = {}
    QF = extract_class(f"{REF}/face_recognition.py", "FaceQualityFilter")
    Aligner = extract_class(f"{REF}/face_recognition.py", "FaceAligner")
    Proc = extract_class(f"{REF}/face_recognition.py", "FaceProcessor", drop=("__init__", "process_image"))

    # ---- quality filter on its own
    rng = np.random.default_rng(321)
    faces = synth_faces(rng, 96)
    distinct = []
    for i in range(20):                                  # 20 distinct crops (4 per sharpness), faces use them in turn
        img = rng.integers(0, 256, (112, 112, 3), dtype=np.uint8)
        sig = [0.0, 0.8, 1.5, 3.0, 6.0][i % 5]
        distinct.append(cv2.GaussianBlur(img, (0, 0), sig) if sig > 0 else img)
    crop_of = rng.integers(0, len(distinct), len(faces))
    crops = [distinct[j] for j in crop_of]
    configs = [dict(), dict(min_det_score=0.6, blur_threshold=100), dict(min_det_score=0.5, min_face_size=40, max_yaw=30, max_pitch=20, max_roll=45, blur_threshold=500),
               dict(check_blur=False)]
    KEYS = ["det_score", "face_size", "yaw", "pitch", "roll", "blur_score"]
    pack["quality/bbox"] = np.stack([f["bbox"] for f in faces])
    pack["quality/landmarks"] = np.stack([f["landmarks"] for f in faces])
    pack["quality/det_score"] = np.array([f["det_score"] for f in faces], np.float32)
    pack["quality/crops"] = np.stack(distinct)
    pack["quality/crop_of"] = crop_of.astype(np.int64)
    pack["quality/configs"] = np.array([repr(c) for c in configs])
    for ci, cfg in enumerate(configs):
        qf = QF(**cfg)
        valid, present, vals = [], [], []
        for f, c in zip(faces, crops):
            ok, m = qf.is_valid(f, c)
            valid.append(ok)
            present.append([k in m for k in KEYS])
            vals.append([float(m.get(k, 0.0)) for k in KEYS])
        pack[f"quality/{ci}/valid"] = np.array(valid)
        pack[f"quality/{ci}/present"] = np.array(present)
        pack[f"quality/{ci}/values"] = np.array(vals, np.float64)

    # ---- process_numpy with a synthetic detector
    rng = np.random.default_rng(654)
    frame = cv2.GaussianBlur(rng.integers(0, 256, (360, 480, 3), dtype=np.uint8), (0, 0), 0.5)
    frame[:, 240:] = cv2.GaussianBlur(frame[:, 240:], (0, 0), 3.0)       # the right half is blurrier
    dets = synth_faces(rng, 14, W=480, H=360)
    pack["process/frame"] = frame
    pack["process/bbox"] = np.stack([f["bbox"] for f in dets])
    pack["process/landmarks"] = np.stack([f["landmarks"] for f in dets])
    pack["process/det_score"] = np.array([f["det_score"] for f in dets], np.float32)
    for S in (112, 224):
        for ci, cfg in enumerate((dict(min_det_score=0.6, blur_threshold=100), dict(min_det_score=0.5, blur_threshold=5, max_roll=60))):
            proc = Proc.__new__(Proc)
            proc.detector = types.SimpleNamespace(detect=lambda img: [dict(f) for f in dets])
            proc.aligner = Aligner(output_size=S)
            proc.quality_filter = QF(**cfg)
            for ra in (True, False):
                res = proc.process_numpy(frame, return_all=ra)
                order = [next(i for i, f in enumerate(dets) if np.array_equal(f["bbox"], r["bbox"]) and np.array_equal(f["landmarks"], r["landmarks"])) for r in res]
                tag = f"process/S{S}/c{ci}/{'all' if ra else 'best'}"
                pack[tag + "/order"] = np.array(order, np.int64)
                pack[tag + "/valid"] = np.array([r["is_valid"] for r in res])
                pack[tag + "/blur"] = np.array([r["quality_metrics"].get("blur_score", -1.0) for r in res], np.float64)
                if ra and S == 112 and ci == 0:
                    pack["process/aligned112"] = np.stack([r["aligned_face"] for r in res])
        pack[f"process/config{0}"] = np.array(repr(dict(min_det_score=0.6, blur_threshold=100)))
        pack[f"process/config{1}"] = np.array(repr(dict(min_det_score=0.5, blur_threshold=5, max_roll=60)))

    # ---- server best-frame selection
    Tracker = extract_class(f"{REF}/face_recognition_server.py", "LiveRecognitionTracker")
    rng = np.random.default_rng(987)
    seg, det, blur, has_q, best, gate = [0], [], [], [], [], []
    for t in range(300):
        tr = Tracker(buffer_size=64)
        F = int(rng.choice([1, 1, 2, 3, 5, 10, 10, 33, 64]))
        for f in range(F):
            d = float(np.float32(rng.choice([0.3, 0.55, 0.6, 0.61, 0.8, 0.8, 0.95]) + (rng.uniform(-0.02, 0.02) if rng.random() < 0.6 else 0.0)))
            b = float(rng.choice([0.0, 20.0, 99.9, 100.0, 150.0, 150.0, 3000.0]) * (rng.uniform(0.9, 1.1) if rng.random() < 0.5 else 1.0))
            q = rng.random() > 0.1
            face = dict(det_score=d, quality_metrics=dict(blur_score=b) if q else {}, frame=f)
            if rng.random() < 0.05:
                face.pop("det_score")                                   # .get('det_score', 0)
                d = 0.0
            tr.add_frame(7, face, "2025-01-01T00:00:00")
            det.append(d); blur.append(b if q else 0.0); has_q.append(q)
        bf = tr.get_best_frame(7)
        best.append(bf["frame"])
        gate.append(bool(tr.should_recognize(7, 0)))
        seg.append(seg[-1] + F)
    pack["bestframe/seg"] = np.array(seg, np.int64)
    pack["bestframe/det"] = np.array(det, np.float64)
    pack["bestframe/blur"] = np.array(blur, np.float64)
    pack["bestframe/best"] = np.array(best, np.int64)
    pack["bestframe/gate"] = np.array(gate)
    np.savez_compressed(os.path.join(OUT, "r2_cases.npz"), **pack)
    print({k: getattr(v, "shape", None) for k, v in pack.items()})


if __name__ == "__main__":
    main()
