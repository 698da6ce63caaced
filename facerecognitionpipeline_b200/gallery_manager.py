"""Drop-in for the reference `gallery_manager` module (gallery_manager.py:16-330).

Same public surface, same pickle / sidecar-JSON / backup layout, same print-and-return-False error
conventions.  What changes underneath:
  * the N x 512 template matrix is built once and kept resident on the GPU (the reference re-runs
    np.vstack over every student on EVERY search call, gallery_manager.py:177-187);
  * `search` runs on the device through the C ABI (`frb_match_host`): query normalisation,
    bf16 tensor-core filter + exact f64 re-score for large galleries, exact scan for small ones,
    canonical tie-break (score desc, insertion index asc — the reference's argsort()[::-1] order on
    ties is unspecified);
  * `search_batch` (new) matches many probes in one call.
Pickles stay interchangeable with the reference: `StudentRecord.__module__` is 'gallery_manager'.
"""
from __future__ import annotations

import json
import os
import pickle
import shutil
import sys
from dataclasses import dataclass
from datetime import datetime
from pathlib import Path
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import templates

SCRIPT_DIR = Path(__file__).resolve().parent


@dataclass
class StudentRecord:
    student_id: str
    name: str
    embeddings: np.ndarray
    template_embedding: np.ndarray
    num_samples: int
    enrollment_date: str
    last_updated: str
    metadata: Dict = None

    def to_dict(self):
        return {
            "student_id": self.student_id,
            "name": self.name,
            "embeddings": self.embeddings.tolist(),
            "template_embedding": self.template_embedding.tolist(),
            "num_samples": self.num_samples,
            "enrollment_date": self.enrollment_date,
            "last_updated": self.last_updated,
            "metadata": self.metadata or {},
        }

    # NOTE: in the reference `from_dict` is dead code (accidentally nested inside to_dict after the
    # return, gallery_manager.py:39-50); it is provided here as the obviously intended classmethod.
    @classmethod
    def from_dict(cls, data: Dict):
        return cls(student_id=data["student_id"], name=data["name"], embeddings=np.array(data["embeddings"]),
                   template_embedding=np.array(data["template_embedding"]), num_samples=data["num_samples"],
                   enrollment_date=data["enrollment_date"], last_updated=data["last_updated"],
                   metadata=data.get("metadata", {}))


# Pickle compatibility in both directions: files written here load in the reference
# (it resolves `gallery_manager.StudentRecord`) and reference pickles load here.
StudentRecord.__module__ = "gallery_manager"
sys.modules.setdefault("gallery_manager", sys.modules[__name__])


class GalleryManager:
    def __init__(self, gallery_path=None, aggregation_method="mean", device: int = 0):
        if gallery_path is None:
            gallery_path = str(SCRIPT_DIR / "gallery" / "students.pkl")
        self.gallery_path = gallery_path
        self.aggregation_method = aggregation_method
        self.students: Dict[str, StudentRecord] = {}
        self._device = device
        self._ctx = None
        self._version = 0          # bumped by every mutation of self.students
        self._resident = None      # (self._version, ctx gallery generation) after this object's last upload
        self._ids: List[str] = []

        os.makedirs(os.path.dirname(gallery_path) or ".", exist_ok=True)
        if os.path.exists(gallery_path):
            self.load()
            print(f"Loaded gallery with {len(self.students)} students")
        else:
            print("Initialized empty gallery")

    # ------------------------------------------------------------------ enrolment (host glue)
    def add_student(self, student_id: str, name: str, embeddings: np.ndarray, metadata: Optional[Dict] = None,
                    overwrite: bool = False) -> bool:
        if student_id in self.students and not overwrite:
            print(f"Student {student_id} already exists. Use overwrite=True to replace.")
            return False
        if embeddings.ndim == 1:
            embeddings = embeddings.reshape(1, -1)
        template = self._aggregate_embeddings(embeddings)
        now = datetime.now().isoformat()
        self.students[student_id] = StudentRecord(student_id=student_id, name=name, embeddings=embeddings,
                                                  template_embedding=template, num_samples=len(embeddings),
                                                  enrollment_date=now, last_updated=now, metadata=metadata or {})
        self._version += 1
        print(f"{'Updated' if overwrite else 'Added'} student: {name} ({student_id}) with {len(embeddings)} embeddings")
        return True

    def _filter_quality_embeddings(self, embeddings: np.ndarray, min_similarity: float = 0.70) -> np.ndarray:
        return templates.quality_filter(embeddings, min_similarity)

    def update_embeddings(self, student_id: str, new_embeddings: np.ndarray, mode: str = "append") -> bool:
        if student_id not in self.students:
            print(f"Student {student_id} not found")
            return False
        student = self.students[student_id]
        if new_embeddings.ndim == 1:
            new_embeddings = new_embeddings.reshape(1, -1)
        if mode == "append":
            updated = np.vstack([student.embeddings, new_embeddings])
        elif mode == "replace":
            updated = new_embeddings
        elif mode == "merge":
            updated = self._remove_outliers(np.vstack([student.embeddings, new_embeddings]))
        else:
            raise ValueError(f"Unknown mode: {mode}")
        student.embeddings = updated
        student.template_embedding = self._aggregate_embeddings(updated)
        student.num_samples = len(updated)
        student.last_updated = datetime.now().isoformat()
        self._version += 1
        print(f"Updated embeddings for {student.name} ({student_id}): {len(student.embeddings)} total embeddings")
        return True

    def delete_student(self, student_id: str) -> bool:
        if student_id not in self.students:
            print(f"Student {student_id} not found")
            return False
        name = self.students[student_id].name
        del self.students[student_id]
        self._version += 1
        print(f"Deleted student: {name} ({student_id})")
        return True

    def get_student(self, student_id: str) -> Optional[StudentRecord]:
        return self.students.get(student_id)

    def get_all_students(self) -> Dict[str, StudentRecord]:
        return self.students

    def get_gallery_embeddings(self) -> Tuple[np.ndarray, List[str]]:
        if len(self.students) == 0:
            return np.array([]), []
        ids = list(self.students.keys())
        return np.vstack([self.students[sid].template_embedding for sid in ids]), ids

    # ------------------------------------------------------------------ device residency + search
    def _context(self):
        if self._ctx is None:
            from . import _native
            self._ctx = _native.default_context(self._device)
        return self._ctx

    def _ensure_resident(self):
        """Upload the template matrix if the student set changed since the last upload, or anything else was
        uploaded to the shared context in between (the ctx counts its uploads: `frb_gallery_generation`).
        Call with the context's lock held."""
        ctx = self._context()
        if self._resident == (self._version, ctx.gallery_generation()):
            return
        mat, ids = self.get_gallery_embeddings()
        mat = np.ascontiguousarray(mat, dtype=np.float32).reshape(len(ids), 512)
        ctx.frb_gallery_upload(mat.ctypes.data, len(ids), 0, 0)
        self._ids = ids
        self._resident = (self._version, ctx.gallery_generation())

    def search_batch(self, query_embeddings: np.ndarray, top_k: int = 5, threshold: float = 0.0):
        """Match P probes at once. Returns (results, accept) where results[p] is the `search` list for
        probe p and accept[p] = (top-1 score >= threshold).  Any top_k >= 1 like the reference
        (gallery_manager.py:197): rows come back as min(top_k, N) tuples."""
        q = np.ascontiguousarray(query_embeddings, dtype=np.float32).reshape(-1, 512)
        P = len(q)
        if len(self.students) == 0 or P == 0:
            return [[] for _ in range(P)], np.zeros(P, dtype=bool)
        k = max(1, min(int(top_k), len(self.students)))
        scores = np.empty((P, k), np.float32)
        idx = np.empty((P, k), np.int64)
        acc = np.empty((P,), np.uint8)
        ctx = self._context()
        with ctx.lock:
            self._ensure_resident()
            ids = self._ids
            ctx.frb_match_host(q.ctypes.data, P, k, float(threshold), 1, scores.ctypes.data, idx.ctypes.data,
                               acc.ctypes.data)
        out = []
        for p in range(P):
            row = []
            for j in range(k):
                gi = int(idx[p, j])
                if gi < 0:
                    break
                sid = ids[gi]
                row.append((sid, self.students[sid].name, float(scores[p, j])))
            out.append(row)
        return out, acc.astype(bool)

    def search(self, query_embedding: np.ndarray, top_k: int = 5) -> List[Tuple[str, str, float]]:
        if len(self.students) == 0:
            return []
        return self.search_batch(np.asarray(query_embedding).reshape(1, -1), top_k)[0][0]

    # ------------------------------------------------------------------ persistence (layout unchanged)
    def save(self, path: Optional[str] = None):
        save_path = path or self.gallery_path
        with open(save_path, "wb") as f:
            pickle.dump(self.students, f)
        json_path = save_path.replace(".pkl", ".json")
        sidecar = {
            "num_students": len(self.students),
            "last_saved": datetime.now().isoformat(),
            "students": {
                sid: {"student_id": s.student_id, "name": s.name, "num_samples": s.num_samples,
                      "enrollment_date": s.enrollment_date, "last_updated": s.last_updated, "metadata": s.metadata}
                for sid, s in self.students.items()
            },
        }
        with open(json_path, "w") as f:
            json.dump(sidecar, f, indent=2)
        print(f"Gallery saved to {save_path}")
        print(f"Metadata saved to {json_path}")

    def load(self, path: Optional[str] = None):
        load_path = path or self.gallery_path
        if not os.path.exists(load_path):
            print(f"Gallery file not found: {load_path}")
            return
        with open(load_path, "rb") as f:
            self.students = pickle.load(f)
        self._version += 1
        print(f"Gallery loaded from {load_path}")

    def export_for_backup(self, backup_dir: str, backup_name: str = None):
        os.makedirs(backup_dir, exist_ok=True)
        timestamp = datetime.now().strftime("%Y%m%d_%H%M%S")
        stem = f"{backup_name}_backup_{timestamp}" if backup_name else f"gallery_backup_{timestamp}"
        backup_path = os.path.join(backup_dir, stem + ".pkl")
        json_path = os.path.join(backup_dir, stem + ".json")
        shutil.copy2(self.gallery_path, backup_path)
        payload = {
            "backup_date": datetime.now().isoformat(),
            "backup_name": backup_name,
            "num_students": len(self.students),
            "students": {sid: s.to_dict() for sid, s in self.students.items()},
        }
        with open(json_path, "w") as f:
            json.dump(payload, f, indent=2)
        print(f"Backup saved to {backup_dir}")

    def get_statistics(self) -> Dict:
        if len(self.students) == 0:
            return {"num_students": 0, "total_embeddings": 0, "avg_embeddings_per_student": 0}
        total = sum(s.num_samples for s in self.students.values())
        return {
            "num_students": len(self.students),
            "total_embeddings": total,
            "avg_embeddings_per_student": total / len(self.students),
            "students": [{"id": s.student_id, "name": s.name, "num_samples": s.num_samples,
                          "enrollment_date": s.enrollment_date} for s in self.students.values()],
        }

    def _aggregate_embeddings(self, embeddings: np.ndarray) -> np.ndarray:
        return templates.gallery_template(embeddings, self.aggregation_method)

    def _remove_outliers(self, embeddings: np.ndarray, threshold: float = 0.7) -> np.ndarray:
        return templates.drop_outliers(embeddings, threshold)
