"""Drop-in for the reference `enroll_students` flow (enroll_students.py:20-486).

What the reference does per student directory (enroll_students.py:119-260): detect + align at 224 px,
keep the best `max_faces` faces, make 8 augmented variants of each (`augment_face_for_enrollment`),
embed them with `FaceEmbedder.extract_embeddings_batch`, measure the mean intra-class similarity from
the gram matrix and hand the embeddings to `GalleryManager.add_student` (weighted-mean template).

Here the same decisions and the same result dictionaries are kept, but the flow is organised for
the device: `enroll_from_directory` first collects the augmented crops of EVERY student on the host
(`collect_student_faces`), embeds them all in device-sized batches through one embedder (eval-mode
embeddings do not depend on the batch a face is in, so this is the per-student result bit for
bit), and only then aggregates per student.  Detection is not part of the hot path (SURVEY §2 row
6): pass `detector=` (an object with `detect(image_rgb) -> [{'bbox','landmarks','det_score'}]`) or a
ready `face_processor=`; the constructor refuses to run without one, as the reference does when
insightface is missing.
"""
from __future__ import annotations

import argparse
import os
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .gallery_manager import GalleryManager

PROJECT_ROOT = Path(__file__).resolve().parent.parent
IMAGE_EXTENSIONS = {".jpg", ".jpeg", ".png", ".bmp"}
AUGMENTATIONS_PER_FACE = 8          # enroll_students.py:210
LOW_INTRA_CLASS_SIMILARITY = 0.3    # enroll_students.py:230


def augment_face_for_enrollment(face_image: np.ndarray, num_augmentations: int = 8) -> List[np.ndarray]:
    """The reference's augmentation list, truncated to `num_augmentations` (enroll_students.py:20-48):
    original, horizontal flip, rotations by -10/-5/+5/+10 degrees about the centre (bilinear,
    replicated border), brightness offsets -20/-10/+10/+20, contrast gains 0.85/0.92/1.08/1.15, a 3x3
    sigma-0.5 blur and additive N(0,3) noise.  With the default of 8 only the first eight are ever
    produced, so the noise variant's unseeded RNG never matters; variants are generated lazily."""
    import cv2
    h, w = face_image.shape[:2]

    def rotated(angle):
        M = cv2.getRotationMatrix2D((w // 2, h // 2), angle, 1.0)
        return cv2.warpAffine(face_image, M, (w, h), borderMode=cv2.BORDER_REPLICATE)

    def shifted(beta):
        return np.clip(face_image.astype(np.float32) + beta, 0, 255).astype(np.uint8)

    def scaled(alpha):
        return np.clip(face_image.astype(np.float32) * alpha, 0, 255).astype(np.uint8)

    def noisy():
        noise = np.random.normal(0, 3, face_image.shape).astype(np.float32)
        return np.clip(face_image.astype(np.float32) + noise, 0, 255).astype(np.uint8)

    makers = [lambda: face_image.copy(), lambda: cv2.flip(face_image, 1)]
    makers += [lambda a=a: rotated(a) for a in (-10, -5, 5, 10)]
    makers += [lambda b=b: shifted(b) for b in (-20, -10, 10, 20)]
    makers += [lambda a=a: scaled(a) for a in (0.85, 0.92, 1.08, 1.15)]
    makers += [lambda: cv2.GaussianBlur(face_image, (3, 3), 0.5), noisy]
    return [make() for make in makers[:max(0, num_augmentations)]]


def intra_class_similarity(embeddings: np.ndarray) -> float:
    """Mean off-diagonal entry of the gram matrix (enroll_students.py:227-228)."""
    n = len(embeddings)
    gram = np.dot(embeddings, embeddings.T)
    return float((np.sum(gram) - n) / (n * (n - 1)))


class StudentEnrollment:
    def __init__(self, gallery_path=PROJECT_ROOT / "gallery" / "students.pkl", min_faces_per_student=3,
                 max_faces_per_student=5, limit_images=0, image_indices=None, model_type="adaface",
                 architecture="ir_101", *, face_processor=None, detector=None, embedder=None, gallery=None,
                 verbose: bool = True, window_crops: int = 2048):
        self.window_crops = int(window_crops)   # crops collected on the host before a window is embedded + registered
        self.min_faces = min_faces_per_student
        self.max_faces = max_faces_per_student
        self.limit_images = limit_images
        self.image_indices = image_indices
        self.verbose = verbose
        if face_processor is None:
            from .face_recognition import FaceProcessor
            face_processor = FaceProcessor(
                output_size=224, det_size=(640, 640), det_thresh=0.5,
                quality_filter_config={"min_det_score": 0.6, "min_face_size": 60, "max_yaw": 45, "max_pitch": 30,
                                       "max_roll": 30, "check_blur": True, "blur_threshold": 100},
                providers=["CUDAExecutionProvider", "CPUExecutionProvider"], detector=detector)
        self.face_processor = face_processor
        if embedder is None:
            from .face_embedder import FaceEmbedder
            embedder = FaceEmbedder(architecture=architecture, model_type=model_type)
        self.embedder = embedder
        self.gallery = gallery if gallery is not None else GalleryManager(gallery_path=str(gallery_path),
                                                                           aggregation_method="weighted_mean")

    def _say(self, *a, **k):
        if self.verbose:
            print(*a, **k)

    # ------------------------------------------------------------------ host half: files -> augmented crops
    def _select_images(self, files: List[str]) -> List[str]:
        if self.image_indices:
            picked = []
            for idx in self.image_indices:
                if 1 <= idx <= len(files):
                    picked.append(files[idx - 1])
                else:
                    self._say(f"Warning: image index {idx} out of range (1-{len(files)})")
            return picked
        if self.limit_images > 0:
            return files[:self.limit_images]
        return files

    def collect_student_faces(self, student_dir: str) -> Tuple[Optional[Dict], Dict]:
        """Detect + align + select + augment for one student.  Returns (work, info): `work` is None when the
        student cannot be enrolled (info then carries the reference's error dictionary), else a dict with the
        augmented crops and the counters `add_student`'s metadata needs."""
        files = sorted(os.path.join(student_dir, f) for f in os.listdir(student_dir)
                       if os.path.splitext(f)[1].lower() in IMAGE_EXTENSIONS)
        if not files:
            self._say(f"No images found in {student_dir}")
            return None, {"error": "no_images"}
        image_files = self._select_images(files)
        seen, valid = 0, []
        for img_path in image_files:
            try:
                faces = self.face_processor.process_image(img_path, return_all=True)
            except Exception as e:  # a bad image must not abort the student (enroll_students.py:186-188)
                self._say(f"  Error processing {os.path.basename(img_path)}: {e}")
                continue
            if not faces:
                continue
            seen += 1
            if faces[0]["is_valid"]:
                valid.append(faces[0])
        self._say(f"  Summary: {len(valid)}/{seen} valid faces")
        if len(valid) < self.min_faces:
            return None, {"error": "insufficient_faces", "valid_faces": len(valid), "required": self.min_faces}
        if len(valid) > self.max_faces:
            valid.sort(key=lambda f: f["det_score"] * f["quality_metrics"].get("blur_score", 1000), reverse=True)
            valid = valid[:self.max_faces]
        crops: List[np.ndarray] = []
        for face in valid:
            crops.extend(augment_face_for_enrollment(face["aligned_face"], num_augmentations=AUGMENTATIONS_PER_FACE))
        return dict(crops=crops, num_images=len(image_files), num_valid_faces=len(valid)), {}

    # ------------------------------------------------------------------ device half + bookkeeping
    def _register(self, student_dir: str, student_id: str, work: Dict, embeddings: np.ndarray) -> Tuple[bool, Dict]:
        name = os.path.basename(student_dir)
        avg = intra_class_similarity(embeddings)
        if avg < LOW_INTRA_CLASS_SIMILARITY:
            self._say(f"Warning: Low intra-class similarity ({avg:.4f}) for {name}")
        ok = self.gallery.add_student(
            student_id=student_id, name=name, embeddings=embeddings,
            metadata={"num_images": work["num_images"], "num_valid_faces": work["num_valid_faces"],
                      "num_augmented_faces": len(work["crops"]), "augmentation_per_face": AUGMENTATIONS_PER_FACE,
                      "avg_similarity": float(avg), "source_directory": student_dir},
            overwrite=True)
        return ok, {"student_id": student_id, "name": name, "num_images": work["num_images"],
                    "num_valid_faces": work["num_valid_faces"], "num_embeddings": len(embeddings),
                    "avg_similarity": float(avg)}

    def _next_student_id(self, offset: int = 0) -> str:
        return f"STU{len(self.gallery.get_all_students()) + 1 + offset:04d}"

    def process_student_directory(self, student_dir: str, student_id: str = None) -> Tuple[bool, Dict]:
        if student_id is None:
            student_id = self._next_student_id()
        self._say(f"Processing: {os.path.basename(student_dir)}  (Student ID: {student_id})")
        work, info = self.collect_student_faces(student_dir)
        if work is None:
            return False, info
        embeddings = self.embedder.extract_embeddings_batch(work["crops"], normalize=True)
        return self._register(student_dir, student_id, work, embeddings)

    def enroll_from_directory(self, enrollment_dir: str) -> Dict:
        if not os.path.exists(enrollment_dir):
            raise ValueError(f"Enrollment directory not found: {enrollment_dir}")
        student_dirs = [os.path.join(enrollment_dir, d) for d in sorted(os.listdir(enrollment_dir))
                        if os.path.isdir(os.path.join(enrollment_dir, d))]
        if not student_dirs:
            self._say("No student directories found!")
            return {"error": "no_directories"}
        # Students are collected on the host and embedded in bounded WINDOWS (about `window_crops` crops, ~150 KB each):
        # the device still sees large batches, host memory no longer grows with the enrollment tree, and every window
        # is registered before the next one is read, so a late failure loses one window, not everything.
        results, successful, failed = [], 0, 0
        window: List[Tuple[str, Optional[Dict], Dict]] = []
        queued = 0

        def flush():
            nonlocal successful, failed, queued
            crops = [c for _, work, _ in window if work is not None for c in work["crops"]]
            emb = self.embedder.extract_embeddings_batch(crops, normalize=True) if crops else np.zeros((0, 512), np.float32)
            at = 0
            for student_dir, work, info in window:
                if work is None:
                    ok = False
                else:
                    n = len(work["crops"])
                    # ids follow the gallery's size at the time the student is added, like the reference's
                    # per-student loop (re-enrolling an existing name overwrites under a NEW id there too)
                    ok, info = self._register(student_dir, self._next_student_id(), work, emb[at:at + n])
                    at += n
                successful += int(ok)
                failed += int(not ok)
                results.append({"directory": student_dir, "success": ok, "info": info})
            window.clear()
            queued = 0

        for d in student_dirs:
            work, info = self.collect_student_faces(d)
            window.append((d, work, info))
            queued += len(work["crops"]) if work is not None else 0
            if queued >= self.window_crops:
                flush()
        flush()
        self.gallery.save()
        stats = self.gallery.get_statistics()
        self._say(f"Total students processed: {len(student_dirs)}  enrolled: {successful}  failed: {failed}")
        if successful > 0:
            self.verify_enrollment()
        return {"total": len(student_dirs), "successful": successful, "failed": failed, "results": results,
                "gallery_stats": stats}

    def verify_enrollment(self) -> Optional[Dict]:
        """Every student's first stored embedding must retrieve that student at rank 1
        (enroll_students.py:350-402); all queries go to the device in one batch."""
        students = self.gallery.get_all_students()
        if len(students) < 2:
            self._say("Need at least 2 students for verification")
            return None
        records = list(students.values())
        queries = np.stack([np.asarray(r.embeddings[0], dtype=np.float32) for r in records])
        hits, _ = self.gallery.search_batch(queries, top_k=3)
        correct, inter = 0, []
        for rec, res in zip(records, hits):
            if res and res[0][1] == rec.name:
                correct += 1
            else:
                self._say(f"  Invalid {rec.name}: matched to {res[0][1] if res else None}")
            inter.extend(score for _, _, score in res[1:])
        total = len(records)
        report = {"rank1_correct": correct, "total": total, "accuracy": correct / total * 100,
                  "avg_inter_class": float(np.mean(inter)) if inter else 0.0,
                  "max_inter_class": float(np.max(inter)) if inter else 0.0}
        self._say(f"Verification: Rank-1 {correct}/{total} ({report['accuracy']:.1f}%), inter-class avg "
                  f"{report['avg_inter_class']:.3f} max {report['max_inter_class']:.3f}")
        if report["max_inter_class"] > 0.6:
            self._say("Warning: high inter-class similarity; some students may be duplicates")
        return report


def main(argv: Optional[Sequence[str]] = None, detector=None):
    ap = argparse.ArgumentParser(description="Enroll students from directory structure")
    ap.add_argument("--enrollment_dir", type=str, default=str(PROJECT_ROOT / "samples" / "enrollment"))
    ap.add_argument("--gallery_path", type=str, default=str(PROJECT_ROOT / "gallery" / "students.pkl"))
    ap.add_argument("--min_faces", type=int, default=1)
    ap.add_argument("--max_faces", type=int, default=5)
    ap.add_argument("--limit_images", type=int, default=0)
    ap.add_argument("--image_indices", type=int, nargs="*", default=None)
    ap.add_argument("--model_type", type=str, default="adaface", choices=["adaface", "arcface"])
    ap.add_argument("--architecture", type=str, default="ir_101", choices=["ir_50", "ir_101"])
    args = ap.parse_args(argv)
    enrollment = StudentEnrollment(gallery_path=args.gallery_path, min_faces_per_student=args.min_faces,
                                   max_faces_per_student=args.max_faces, limit_images=args.limit_images,
                                   image_indices=args.image_indices, model_type=args.model_type,
                                   architecture=args.architecture, detector=detector)
    summary = enrollment.enroll_from_directory(args.enrollment_dir)
    if summary.get("successful", 0) > 0:
        backup_dir = Path(args.gallery_path).parent / "backups"
        enrollment.gallery.export_for_backup(str(backup_dir), backup_name=f"{args.model_type}_{args.architecture}")
    return summary


if __name__ == "__main__":
    main()
