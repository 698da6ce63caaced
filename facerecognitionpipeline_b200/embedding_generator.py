"""Drop-in for the reference `embedding_generator` flow (embedding_generator.py:54-506): bulk embedding
of gallery / positive-probe / negative-probe image trees for one model configuration, written as
`<output_root>/embeddings/<model_type>_<arch>/{gallery_<type>_<base|augmented>, probe_positive_
<segmented|unsegmented>, probe_negative}.{pkl,json}` + `generation_summary.json`, same keys and shapes.

Device-first differences (results are the reference's, only the batching changes):
  * the reference embeds pre-cropped probes ONE image per forward pass (embedding_generator.py:268,330)
    and forces batch 1 for ArcFace (:190); here every directory becomes one list handed to
    `extract_embeddings_batch`, which streams device-sized batches (an eval-mode embedding does not
    depend on the batch it is computed in);
  * detection is outside the hot path: pass `detector=` / `face_processor=` for the gallery trees;
    probe trees hold pre-aligned crops and need none.
"""
from __future__ import annotations

import argparse
import json
import pickle
from datetime import datetime
from pathlib import Path
from typing import Dict, List, Optional, Sequence

import numpy as np

from .enroll_students import augment_face_for_enrollment

PROJECT_ROOT = Path(__file__).resolve().parent.parent
SEGMENT_CATEGORIES = ["high_quality", "blur_blurry", "blur_sharp", "face_large", "face_medium", "face_small",
                      "pose_easy", "pose_medium", "pose_hard", "low_quality"]   # embedding_generator.py:223-224


def _image_files(directory: Path) -> List[Path]:
    # the reference concatenates three globs and sorts the result (embedding_generator.py:155-157)
    return sorted(list(directory.glob("*.jpg")) + list(directory.glob("*.png")) + list(directory.glob("*.jpeg")))


def to_serializable(obj):
    """ndarray / numpy scalar -> plain Python, recursively (embedding_generator.py:109-122)."""
    if isinstance(obj, np.ndarray):
        return obj.tolist()
    if isinstance(obj, dict):
        return {k: to_serializable(v) for k, v in obj.items()}
    if isinstance(obj, list):
        return [to_serializable(v) for v in obj]
    if isinstance(obj, (np.int64, np.int32)):
        return int(obj)
    if isinstance(obj, (np.float64, np.float32)):
        return float(obj)
    return obj


class EmbeddingGenerator:
    def __init__(self, model_type="adaface", architecture="ir_101", dataset_root=None, output_root=None, *,
                 embedder=None, face_processor=None, detector=None, verbose: bool = True):
        self.model_type = model_type
        self.architecture = architecture
        self.model_name = f"{model_type}_{architecture}"
        self.dataset_root = Path(dataset_root) if dataset_root is not None else PROJECT_ROOT / "dataset"
        self.output_root = Path(output_root) if output_root is not None else PROJECT_ROOT / "output" / "v0"
        self.verbose = verbose
        if embedder is None:
            from .face_embedder import FaceEmbedder
            embedder = FaceEmbedder(architecture=architecture, model_type=model_type)
        self.embedder = embedder
        if face_processor is None and detector is not None:
            from .face_recognition import FaceProcessor
            face_processor = FaceProcessor(output_size=112, det_size=(640, 640), det_thresh=0.5,
                                           quality_filter_config={"min_det_score": 0.5, "min_face_size": 40},
                                           providers=["CUDAExecutionProvider", "CPUExecutionProvider"], detector=detector)
        self.face_processor = face_processor   # only the gallery trees (raw photos) need it
        self.output_dir = self.output_root / "embeddings" / self.model_name
        self.output_dir.mkdir(parents=True, exist_ok=True)

    def _say(self, *a):
        if self.verbose:
            print(*a)

    @staticmethod
    def extract_name_from_filename(filename: str) -> str:
        """'first_last_012_x.jpg' -> 'first_last': underscore-separated parts up to the first all-digit one."""
        parts = Path(filename).stem.split("_")
        keep = []
        for part in parts:
            if part.isdigit():
                break
            keep.append(part)
        return "_".join(keep) if keep else parts[0]

    def save_embeddings_json(self, data: Dict, output_path: Path):
        with open(Path(output_path).with_suffix(".json"), "w") as f:
            json.dump(to_serializable(data), f, indent=2)

    def _dump(self, data: Dict, stem: str) -> Path:
        path = self.output_dir / f"{stem}.pkl"
        with open(path, "wb") as f:
            pickle.dump(data, f)
        self.save_embeddings_json(data, path)
        self._say(f"Saved to: {path}")
        return path

    def load_image(self, image_path: Path) -> np.ndarray:
        import cv2
        img = cv2.imread(str(image_path))
        if img is None:
            raise ValueError(f"Failed to load image: {image_path}")
        return cv2.cvtColor(img, cv2.COLOR_BGR2RGB)

    def _load_crop(self, image_path: Path) -> np.ndarray:
        import cv2
        img = self.load_image(image_path)
        if img.shape[0] != 112 or img.shape[1] != 112:
            img = cv2.resize(img, (112, 112))    # default bilinear, embedding_generator.py:265-266
        return img

    # ------------------------------------------------------------------ gallery trees (raw photos)
    def process_gallery_enrollment(self, enrollment_type: str = "one-shot", use_augmentation: bool = False) -> Dict:
        suffix = "augmented" if use_augmentation else "base"
        gallery_dir = self.dataset_root / "enrollment" / enrollment_type
        if not gallery_dir.exists():
            self._say(f"Warning: Gallery directory not found: {gallery_dir}")
            return {}
        if self.face_processor is None:
            raise ImportError("process_gallery_enrollment needs a face detector (pass detector= or face_processor=)")
        # host pass: per person, the aligned (and optionally augmented) crops; then one device stream
        people = []
        for person_dir in sorted(d for d in gallery_dir.iterdir() if d.is_dir()):
            crops, valid_files = [], []
            for img_path in _image_files(person_dir):
                try:
                    faces = self.face_processor.process_image(str(img_path), return_all=True)
                    if not faces:
                        continue
                    aligned = faces[0]["aligned_face"]
                    crops.extend(augment_face_for_enrollment(aligned, num_augmentations=8) if use_augmentation else [aligned])
                    valid_files.append(img_path.name)
                except Exception as e:  # one bad image must not abort the person (embedding_generator.py:184-186)
                    self._say(f"Error processing {img_path}: {e}")
            if crops:
                people.append((person_dir.name, crops, valid_files))
        all_crops = [c for _, crops, _ in people for c in crops]
        all_emb = self.embedder.extract_embeddings_batch(all_crops, normalize=True) if all_crops else np.zeros((0, 512), np.float32)
        out, at = {}, 0
        for name, crops, valid_files in people:
            emb = all_emb[at:at + len(crops)]
            at += len(crops)
            out[name] = {"embeddings": emb, "num_images": len(valid_files), "num_embeddings": len(emb),
                         "image_files": valid_files, "enrollment_type": enrollment_type, "augmented": use_augmentation}
        self._dump(out, f"gallery_{enrollment_type}_{suffix}")
        return out

    # ------------------------------------------------------------------ probe trees (pre-aligned crops)
    def _embed_files(self, files: Sequence[Path]):
        """Load every crop that loads, embed them in one call; returns (kept files, [n,512] embeddings)."""
        kept, crops = [], []
        for p in files:
            try:
                crops.append(self._load_crop(p))
                kept.append(p)
            except Exception as e:
                self._say(f"Error processing {p.name}: {e}")
        emb = self.embedder.extract_embeddings_batch(crops, normalize=True) if crops else np.zeros((0, 512), np.float32)
        return kept, emb

    def process_probe_positive(self, segmented: bool = False) -> Dict:
        base = self.output_root / "probe_labeled" / ("segmented" if segmented else "positive")
        categories = SEGMENT_CATEGORIES if segmented else ["."]
        if not base.exists():
            self._say(f"Warning: Probe directory not found: {base}")
            return {}
        out = {}
        for category in categories:
            cdir, cname = (base, "all") if category == "." else (base / category, category)
            if not cdir.exists():
                continue
            files = _image_files(cdir)
            if not files:
                continue
            kept, emb = self._embed_files(files)
            per_person: Dict[str, Dict] = {}
            for p, e in zip(kept, emb):
                slot = per_person.setdefault(self.extract_name_from_filename(p.name), {"embeddings": [], "filenames": []})
                slot["embeddings"].append(e)
                slot["filenames"].append(p.name)
            for slot in per_person.values():
                slot["embeddings"] = np.array(slot["embeddings"])
            out[cname] = per_person
        self._dump(out, f"probe_positive_{'segmented' if segmented else 'unsegmented'}")
        return out

    def process_probe_negative(self) -> Dict:
        probe_dir = self.output_root / "probe_labeled" / "negative"
        if not probe_dir.exists():
            self._say(f"Warning: Probe directory not found: {probe_dir}")
            return {}
        out = {"real": {"embeddings": [], "filenames": []}, "lfw": {"embeddings": [], "filenames": []}}
        kept, emb = self._embed_files(_image_files(probe_dir))
        for p, e in zip(kept, emb):
            cat = "lfw" if ("lfw" in p.name.lower() or "lfw" in str(p.parent).lower()) else "real"
            out[cat]["embeddings"].append(e)
            out[cat]["filenames"].append(p.name)
        for cat in out:
            if len(out[cat]["embeddings"]) > 0:
                out[cat]["embeddings"] = np.array(out[cat]["embeddings"])
        self._dump(out, "probe_negative")
        return out

    def generate_all_embeddings(self) -> Dict:
        t0 = datetime.now()
        galleries = {}
        for etype, aug in (("one-shot", False), ("one-shot", True), ("few-shot", False), ("few-shot", True)):
            galleries[(etype, aug)] = self.process_gallery_enrollment(etype, use_augmentation=aug) \
                if (self.dataset_root / "enrollment" / etype).exists() else {}
        pos_unseg = self.process_probe_positive(segmented=False)
        pos_seg = self.process_probe_positive(segmented=True)
        neg = self.process_probe_negative()
        summary = {
            "model_type": self.model_type, "architecture": self.architecture, "model_name": self.model_name,
            "timestamp": datetime.now().isoformat(), "duration_seconds": (datetime.now() - t0).total_seconds(),
            "gallery": {"one_shot_base_persons": len(galleries[("one-shot", False)]),
                        "one_shot_augmented_persons": len(galleries[("one-shot", True)]),
                        "few_shot_base_persons": len(galleries[("few-shot", False)]),
                        "few_shot_augmented_persons": len(galleries[("few-shot", True)])},
            "probe_positive": {"unsegmented_categories": list(pos_unseg.keys()) if pos_unseg else [],
                               "segmented_categories": list(pos_seg.keys()) if pos_seg else []},
            "probe_negative": {"real_images": len(neg.get("real", {}).get("embeddings", [])),
                               "lfw_images": len(neg.get("lfw", {}).get("embeddings", []))},
            "output_directory": str(self.output_dir),
        }
        with open(self.output_dir / "generation_summary.json", "w") as f:
            json.dump(summary, f, indent=2)
        return summary


def main(argv: Optional[Sequence[str]] = None, detector=None):
    ap = argparse.ArgumentParser(description="Generate face embeddings for evaluation using multiple models")
    ap.add_argument("--model_type", type=str, default="all", choices=["adaface", "arcface", "all"])
    ap.add_argument("--architecture", type=str, default="all", choices=["ir_50", "ir_101", "all"])
    ap.add_argument("--dataset_root", type=str, default=None)
    ap.add_argument("--output_root", type=str, default=None)
    args = ap.parse_args(argv)
    model_types = ["adaface", "arcface"] if args.model_type == "all" else [args.model_type]
    architectures = ["ir_50", "ir_101"] if args.architecture == "all" else [args.architecture]
    done = []
    for model_type in model_types:
        for architecture in architectures:
            try:   # one failing configuration must not stop the others (embedding_generator.py:487-497)
                gen = EmbeddingGenerator(model_type=model_type, architecture=architecture, dataset_root=args.dataset_root,
                                         output_root=args.output_root, detector=detector)
                done.append(gen.generate_all_embeddings())
            except Exception as e:
                print(f"ERROR in {model_type}_{architecture}: {e}")
    return done


if __name__ == "__main__":
    main()
