"""Drop-in for the reference `face_matcher.FaceMatcher` (face_matcher.py:19-385).

`match_single_face` (face_matcher.py:52-58) is the per-face unit of the hot path: embed -> search.
Here both halves run on the B200 in one C-ABI call (`frb_embed_match_host`), and
`match_faces_batch` (new) does it for a whole list of crops at once.  The multi-frame consensus
(`_aggregate_matches`, face_matcher.py:321-363) and result dictionaries are host glue with the
reference's exact decision rules and field names; `consensus_on_device` / `match_tracks_batch` run the
same decision for many tracks in one launch (`frb_track_consensus`).
"""
from __future__ import annotations

import json
import os
from collections import Counter
from datetime import datetime
from pathlib import Path
from typing import Dict, List, Optional, Tuple

import numpy as np

from .face_embedder import FaceEmbedder
from .gallery_manager import GalleryManager

SCRIPT_DIR = Path(__file__).resolve().parent

MIN_QUALITY_SCORE = 0.55   # face_matcher.py:324,366
MIN_QUALITY_FRAMES = 3     # face_matcher.py:325


def consensus(frame_matches: List[Dict], similarity_threshold: float) -> Optional[Dict]:
    """Track-level decision of FaceMatcher._aggregate_matches: frames scoring >= 0.55 vote; at least 3
    of them; the winner needs > 50 % of the votes, or > 40 % and at least twice the runner-up; the
    mean of the winner's frame scores must reach the similarity threshold."""
    voters = [m for m in frame_matches if m["score"] >= MIN_QUALITY_SCORE]
    if len(voters) < MIN_QUALITY_FRAMES:
        return None
    ranking = Counter(m["student_id"] for m in voters).most_common(2)
    winner, votes = ranking[0]
    share = votes / len(voters)
    agreed = share > 0.5
    if not agreed and len(ranking) > 1:
        agreed = share > 0.4 and votes >= 2 * ranking[1][1]
    if not agreed:
        return None
    winner_scores = [m["score"] for m in voters if m["student_id"] == winner]
    mean_score = np.mean(winner_scores)
    if mean_score < similarity_threshold:
        return None
    return {
        "student_id": winner,
        "name": next(m["name"] for m in voters if m["student_id"] == winner),
        "confidence": float(mean_score),
        "consensus_strength": float(share),
        "num_quality_frames": len(winner_scores),
        "total_frames_evaluated": len(frame_matches),
    }


def best_candidate(frame_matches: List[Dict]) -> Dict:
    """FaceMatcher._get_best_candidate (face_matcher.py:365-385)."""
    pool = [m for m in frame_matches if m["score"] >= MIN_QUALITY_SCORE] or frame_matches
    sid = Counter(m["student_id"] for m in pool).most_common(1)[0][0]
    scores = [m["score"] for m in pool if m["student_id"] == sid]
    return {
        "student_id": sid,
        "name": next(m["name"] for m in pool if m["student_id"] == sid),
        "confidence": float(np.mean(scores)),
        "num_quality_frames": len(scores),
    }


def consensus_on_device(top_idx, top_score, frames_per_track, similarity_threshold: float, device: int = 0):
    """`consensus` + `best_candidate` for many tracks in one launch (`frb_track_consensus`).

    top_idx / top_score: [F] or [F][k] (column 0 = top-1) gallery rows (int64, < 0 = frame without a match) and f32
    scores of every frame of every track, numpy or CUDA torch tensors (e.g. straight from frb_match);
    frames_per_track: the T track lengths.  Returns a numpy record array with the fields of `frb_track_result`."""
    import torch

    from . import _native
    ctx = _native.default_context(device)
    dev = torch.device("cuda", device)
    ix = torch.as_tensor(top_idx).to(device=dev, dtype=torch.int64).contiguous()
    sc = torch.as_tensor(top_score).to(device=dev, dtype=torch.float32).contiguous()
    stride = 1 if ix.dim() == 1 else int(ix.shape[1])
    if ix.shape != sc.shape:
        raise ValueError("top_idx and top_score differ in shape")
    counts = np.asarray(frames_per_track, dtype=np.int64).reshape(-1)
    T = len(counts)
    if int(counts.sum()) != int(ix.shape[0]):
        raise ValueError(f"frames_per_track sums to {int(counts.sum())} but {int(ix.shape[0])} frames were given")
    dt = np.dtype([("winner", "<i8"), ("confidence", "<f8"), ("consensus_strength", "<f8"), ("num_quality_frames", "<i4"),
                   ("total_frames_evaluated", "<i4"), ("candidate", "<i8"), ("candidate_confidence", "<f8"),
                   ("candidate_num_quality_frames", "<i4"), ("recognized", "<i4")])
    assert dt.itemsize == 56
    if T == 0:
        return np.zeros(0, dt)
    seg = torch.from_numpy(np.concatenate([[0], np.cumsum(counts)])).to(dev)
    out = torch.empty((T, dt.itemsize), dtype=torch.uint8, device=dev)
    ctx.frb_track_consensus(ix.data_ptr(), sc.data_ptr(), stride, seg.data_ptr(), T, int(counts.max()), MIN_QUALITY_SCORE,
                            MIN_QUALITY_FRAMES, float(similarity_threshold), out.data_ptr(),
                            torch.cuda.current_stream(dev).cuda_stream)
    return out.cpu().numpy().view(dt).reshape(T)


BEST_FRAME_MIN_DET = 0.6   # face_recognition_server.py:61: the best frame's detection score must exceed this


def best_frame_index(det_scores, blur_scores) -> Tuple[int, bool]:
    """Server best-frame selection for ONE track (LiveRecognitionTracker.get_best_frame / should_recognize,
    face_recognition_server.py:39-85): quality = det_score * min(blur_score / 100, 1); the FIRST frame with the
    largest quality wins (Python `max`); the track is ready for recognition when that frame's det_score > 0.6.
    A frame without quality metrics counts with blur_score 0."""
    best, best_q = -1, 0.0
    for i, (d, b) in enumerate(zip(det_scores, blur_scores)):
        q = float(d) * min(float(b) / 100.0, 1.0)
        if best < 0 or q > best_q:
            best, best_q = i, q
    return best, bool(best >= 0 and float(det_scores[best]) > BEST_FRAME_MIN_DET)


def best_frames_batch(det_scores, blur_scores, seg, device: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """Same rule for T tracks in one launch (`frb_best_frames`): track t owns frames [seg[t], seg[t+1]).
    Returns (index of the best frame inside its track [T] i64 (-1 = empty track), ready [T] bool)."""
    import ctypes as C

    import torch

    from . import _native
    seg = np.ascontiguousarray(seg, dtype=np.int64)
    T = len(seg) - 1
    if T <= 0:
        return np.zeros((0,), np.int64), np.zeros((0,), bool)
    dev = torch.device("cuda", device)
    d = torch.from_numpy(np.ascontiguousarray(det_scores, dtype=np.float64)).to(dev)
    b = torch.from_numpy(np.ascontiguousarray(blur_scores, dtype=np.float64)).to(dev)
    s = torch.from_numpy(seg).to(dev)
    idx = torch.empty((T,), dtype=torch.int64, device=dev)
    ready = torch.empty((T,), dtype=torch.uint8, device=dev)
    ctx = _native.default_context(device)
    ctx.frb_best_frames(d.data_ptr(), b.data_ptr(), s.data_ptr(), T, BEST_FRAME_MIN_DET, idx.data_ptr(), None,
                        ready.data_ptr(), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    return idx.cpu().numpy(), ready.cpu().numpy().astype(bool)


class FaceMatcher:
    def __init__(self, gallery_path=None, similarity_threshold=0.5, aggregation_method="majority_vote",
                 model_type="adaface", architecture="ir_101", embedder: Optional[FaceEmbedder] = None,
                 gallery: Optional[GalleryManager] = None):
        self.similarity_threshold = similarity_threshold
        self.aggregation_method = aggregation_method
        self.model_type = model_type
        self.architecture = architecture
        if gallery_path is None:
            gallery_path = str(SCRIPT_DIR / "gallery" / "students.pkl")
        print("Initializing Face Matcher...")
        self.embedder = embedder if embedder is not None else FaceEmbedder(architecture=architecture, model_type=model_type)
        self.gallery = gallery if gallery is not None else GalleryManager(gallery_path=gallery_path)
        n = len(self.gallery.get_all_students())
        if n == 0:
            print("\nWARNING: Gallery is empty! Please enroll students first.")
        else:
            print(f"   Loaded {n} enrolled students")
        print("Face Matcher ready!")

    # ------------------------------------------------------------------ hot path
    def match_faces_batch(self, face_images: List[np.ndarray], top_k: int = 5):
        """Embed + match a list of aligned RGB crops.  Returns (results, accept): results[i] is what
        match_single_face(face_images[i]) returns, accept[i] = top-1 score >= similarity_threshold."""
        if len(face_images) == 0:
            return [], np.zeros(0, dtype=bool)
        embeddings = self.embedder.extract_embeddings_batch(face_images, normalize=True)
        return self.gallery.search_batch(embeddings, top_k=top_k, threshold=self.similarity_threshold)

    def match_single_face(self, face_image: np.ndarray, top_k: int = 5) -> List[Tuple[str, str, float]]:
        embedding = self.embedder.extract_embedding(face_image, normalize=True)
        return self.gallery.search(embedding, top_k=top_k)

    # ------------------------------------------------------------------ track flow
    def match_track(self, track_dir: str, top_k: int = 3) -> Optional[Dict]:
        import cv2
        track_id = os.path.basename(track_dir)
        meta_path = os.path.join(track_dir, "metadata.json")
        if not os.path.exists(meta_path):
            print(f"No metadata found for {track_id}")
            return None
        with open(meta_path, "r") as f:
            metadata = json.load(f)
        files = sorted(f for f in os.listdir(track_dir) if f.endswith(".jpg"))
        if not files:
            print(f"No face images found in {track_id}")
            return None
        print(f"\nProcessing {track_id}: {len(files)} frames")
        crops, names = [], []
        for fn in files:
            bgr = cv2.imread(os.path.join(track_dir, fn))
            if bgr is None:
                continue
            crops.append(cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB))
            names.append(fn)
        results, _ = self.match_faces_batch(crops, top_k=top_k)
        frame_matches = []
        for fn, matches in zip(names, results):
            if not matches:
                continue
            sid, name, score = matches[0]
            frame_matches.append({
                "frame": fn, "student_id": sid, "name": name, "score": float(score),
                "top_k_matches": [{"student_id": s, "name": n, "score": float(v)} for s, n, v in matches],
            })
        if not frame_matches:
            print("No valid matches found")
            return None
        all_scores: Dict[str, List[float]] = {}
        for m in frame_matches:
            all_scores.setdefault(m["student_id"], []).append(m["score"])
        final = self._aggregate_matches(frame_matches, all_scores)
        if final is None:
            cand = self._get_best_candidate(frame_matches, all_scores)
            print(f"Below threshold - Best candidate: {cand['name']} ({cand['student_id']}) - confidence: {cand['confidence']:.3f}")
            return {"track_id": track_id, "recognized": False, "reason": "below_threshold", "best_candidate": cand,
                    "frame_matches": frame_matches, "metadata": metadata, "timestamp": datetime.now().isoformat()}
        print(f"  Identified: {final['name']} ({final['student_id']}) - confidence: {final['confidence']:.3f}")
        return {"track_id": track_id, "recognized": True, "student_id": final["student_id"], "name": final["name"],
                "confidence": final["confidence"], "method": self.aggregation_method, "num_frames": len(frame_matches),
                "frame_matches": frame_matches, "metadata": metadata, "timestamp": datetime.now().isoformat()}

    def match_tracks_batch(self, tracks: List[List[np.ndarray]], top_k: int = 3) -> List[Optional[Dict]]:
        """The decision part of `match_track` for many tracks at once: every crop of every track goes through ONE
        embed stream, one device match and one consensus launch; results come back in a single copy.  Entry t is
        None for a track without any matched frame, else the `match_track` dictionary fields that do not depend on
        files (recognized / student_id / name / confidence / ... / best_candidate / frame top-1 list)."""
        import torch
        counts = [len(t) for t in tracks]
        crops = [c for t in tracks for c in t]
        if not crops:
            return [None] * len(tracks)
        if len(self.gallery.students) == 0:
            return [None] * len(tracks)
        emb = self.embedder.extract_embeddings_batch(crops, normalize=True)
        self.gallery._ensure_resident()
        ctx, dev = self.gallery._context(), torch.device("cuda", self.gallery._device)
        F, k = len(crops), int(top_k)
        d_q = torch.from_numpy(np.ascontiguousarray(emb, dtype=np.float32)).to(dev)
        d_sc = torch.empty((F, k), dtype=torch.float32, device=dev)
        d_ix = torch.empty((F, k), dtype=torch.int64, device=dev)
        d_ac = torch.empty((F,), dtype=torch.uint8, device=dev)
        ctx.frb_match(d_q.data_ptr(), F, k, float(self.similarity_threshold), 1, d_sc.data_ptr(), d_ix.data_ptr(),
                      d_ac.data_ptr(), None, torch.cuda.current_stream(dev).cuda_stream)
        res = consensus_on_device(d_ix, d_sc, counts, self.similarity_threshold, self.gallery._device)
        ix, sc = d_ix.cpu().numpy(), d_sc.cpu().numpy()
        ids = self.gallery._ids

        def ident(row):
            sid = ids[int(row)]
            return sid, self.gallery.students[sid].name

        out, f0 = [], 0
        for t, n in enumerate(counts):
            r = res[t]
            if r["total_frames_evaluated"] == 0:
                out.append(None)
                f0 += n
                continue
            frames = [{"student_id": ident(ix[f, 0])[0], "name": ident(ix[f, 0])[1], "score": float(sc[f, 0])}
                      for f in range(f0, f0 + n) if ix[f, 0] >= 0]
            f0 += n
            if r["recognized"]:
                sid, name = ident(r["winner"])
                out.append({"recognized": True, "student_id": sid, "name": name, "confidence": float(r["confidence"]),
                            "consensus_strength": float(r["consensus_strength"]),
                            "num_quality_frames": int(r["num_quality_frames"]), "method": self.aggregation_method,
                            "num_frames": int(r["total_frames_evaluated"]), "frame_matches": frames})
            else:
                sid, name = ident(r["candidate"])
                out.append({"recognized": False, "reason": "below_threshold",
                            "best_candidate": {"student_id": sid, "name": name,
                                               "confidence": float(r["candidate_confidence"]),
                                               "num_quality_frames": int(r["candidate_num_quality_frames"])},
                            "frame_matches": frames})
        return out

    def _aggregate_matches(self, frame_matches: List[Dict], all_scores: Dict[str, List[float]]) -> Optional[Dict]:
        return consensus(frame_matches, self.similarity_threshold)

    def _get_best_candidate(self, frame_matches, all_scores):
        return best_candidate(frame_matches)

    # ------------------------------------------------------------------ single image flow
    def match_single_image(self, image_path: str, top_k: int = 5, save_visualization: bool = True) -> Dict:
        if not os.path.exists(image_path):
            raise ValueError(f"Image not found: {image_path}")
        from .face_recognition import FaceProcessor  # needs insightface for detection (out of scope here)
        processor = FaceProcessor(output_size=112, det_size=(640, 640), det_thresh=0.5, quality_filter_config={
            "min_det_score": 0.5, "min_face_size": 40, "max_yaw": 60, "max_pitch": 45, "max_roll": 45,
            "check_blur": True, "blur_threshold": 50}, providers=["CUDAExecutionProvider", "CPUExecutionProvider"])
        faces = processor.process_image(image_path, return_all=True)
        if not faces:
            print("No faces detected in image")
            return {"image_path": image_path, "num_faces": 0, "matches": [], "timestamp": datetime.now().isoformat()}
        results, accept = self.match_faces_batch([f["aligned_face"] for f in faces], top_k=top_k)
        matches = []
        for idx, (face, res, ok) in enumerate(zip(faces, results, accept)):
            if not res:
                matches.append({"face_index": idx, "bbox": face["bbox"].tolist(), "recognized": False,
                                "reason": "no_gallery_matches", "quality_metrics": face["quality_metrics"]})
                continue
            sid, name, score = res[0]
            recognized = bool(score >= self.similarity_threshold)
            entry = {"face_index": idx, "bbox": face["bbox"].tolist(), "recognized": recognized,
                     "confidence": float(score), "quality_metrics": face["quality_metrics"],
                     "top_matches": [{"student_id": s, "name": n, "score": float(v), "rank": r + 1}
                                     for r, (s, n, v) in enumerate(res)]}
            if recognized:
                entry["student_id"], entry["name"] = sid, name
            else:
                entry["best_candidate"] = {"student_id": sid, "name": name, "confidence": float(score)}
            matches.append(entry)
        return {"image_path": image_path, "num_faces": len(faces),
                "num_recognized": sum(1 for m in matches if m.get("recognized", False)), "matches": matches,
                "threshold": self.similarity_threshold, "timestamp": datetime.now().isoformat()}

    # ------------------------------------------------------------------ directory flow
    def process_capture_directory(self, capture_dir: str, save_results: bool = True) -> Dict:
        if not os.path.exists(capture_dir):
            raise ValueError(f"Capture directory not found: {capture_dir}")
        track_dirs = [os.path.join(capture_dir, d) for d in sorted(os.listdir(capture_dir))
                      if d.startswith("track_") and os.path.isdir(os.path.join(capture_dir, d))]
        if not track_dirs:
            print("No track directories found!")
            return {"error": "no_tracks"}
        results = []
        for td in track_dirs:
            r = self.match_track(td, top_k=3)
            if r is None:
                continue
            results.append(r)
            if save_results:
                with open(os.path.join(td, "recognition_result.json"), "w") as f:
                    json.dump(r, f, indent=2)
        recognized = sum(1 for r in results if r["recognized"])
        summary = self._generate_summary(results, recognized, len(results) - recognized)
        if save_results:
            out_dir = os.path.join(capture_dir, f"{self.model_type}_{self.architecture}_results")
            os.makedirs(out_dir, exist_ok=True)
            with open(os.path.join(out_dir, "recognition_summary.json"), "w") as f:
                json.dump(summary, f, indent=2)
        return summary

    def _generate_summary(self, results: List[Dict], recognized: int, unrecognized: int) -> Dict:
        appearances = Counter(r["name"] for r in results if r["recognized"])
        conf = [r["confidence"] for r in results if r["recognized"]]
        return {
            "total_tracks": len(results),
            "recognized": recognized,
            "unrecognized": unrecognized,
            "recognition_rate": recognized / len(results) * 100 if results else 0,
            "avg_confidence": float(np.mean(conf)) if conf else 0.0,
            "student_appearances": dict(appearances.most_common()),
            "below_threshold_candidates": [r["best_candidate"] for r in results
                                           if not r["recognized"] and "best_candidate" in r],
            "unique_students": len(appearances),
            "timestamp": datetime.now().isoformat(),
            "settings": {"similarity_threshold": self.similarity_threshold,
                         "aggregation_method": self.aggregation_method},
        }
