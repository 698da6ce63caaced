"""Memory-mappable gallery format for large galleries.

The reference persists a gallery as `pickle.dump(Dict[str, StudentRecord])` + a sidecar JSON
(gallery_manager.py:207-232) and backs it up as `.tolist()` JSON (:246-270).  That layout is kept
untouched by `GalleryManager` (drop-in), but it does not scale to the 1M / 10M-identity configurations:
a pickle of 10M dataclasses cannot be opened partially and every `search` rebuilds the matrix with
`np.vstack` (:177-187).  This module adds a flat, shardable companion layout next to it:

    <stem>.templates.npy   [N, 512] float32, numpy .npy (np.load(mmap_mode='r') maps it without reading it)
    <stem>.ids.npy         [N] fixed-width UTF-8 bytes (student ids, row order = template order)
    <stem>.names.npy       [N] fixed-width UTF-8 bytes (display names)
    <stem>.meta.json       {"format": "frb-gallery-matrix-1", "num_students": N, "dim": 512, "source": ...}

Row order is `GalleryManager.get_gallery_embeddings()` order (dict insertion order), so row r of the
matrix is exactly what `search` scores as index r.  `MatrixGallery` maps the files, uploads only the rows
of its identity shard to the device (chunked through pinned memory) and answers `search_batch` with the
same tuples as `GalleryManager.search`.
"""
from __future__ import annotations

import json
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np

FORMAT = "frb-gallery-matrix-1"


def _paths(stem: str):
    return stem + ".templates.npy", stem + ".ids.npy", stem + ".names.npy", stem + ".meta.json"


def write_matrix(stem: str, templates: np.ndarray, ids: Sequence[str], names: Optional[Sequence[str]] = None,
                 source: str = "") -> None:
    templates = np.ascontiguousarray(templates, dtype=np.float32).reshape(len(ids), 512)
    names = list(names) if names is not None else [""] * len(ids)
    if len(names) != len(ids):
        raise ValueError("ids and names differ in length")
    t_path, i_path, n_path, m_path = _paths(stem)
    os.makedirs(os.path.dirname(stem) or ".", exist_ok=True)
    np.save(t_path, templates)
    np.save(i_path, np.array([s.encode("utf-8") for s in ids], dtype=np.bytes_) if len(ids) else np.zeros((0,), "S1"))
    np.save(n_path, np.array([s.encode("utf-8") for s in names], dtype=np.bytes_) if len(ids) else np.zeros((0,), "S1"))
    with open(m_path, "w") as f:
        json.dump({"format": FORMAT, "num_students": len(ids), "dim": 512, "source": source}, f, indent=2)


def export_gallery(gallery, stem: str) -> None:
    """Write the companion layout of a GalleryManager (templates exactly as `search` uses them)."""
    mat, ids = gallery.get_gallery_embeddings()
    names = [gallery.students[sid].name for sid in ids]
    write_matrix(stem, np.asarray(mat, dtype=np.float32).reshape(len(ids), 512), ids, names,
                 source=str(getattr(gallery, "gallery_path", "")))


class MatrixGallery:
    """Read-only, shardable view of a matrix gallery with device-resident search."""

    def __init__(self, stem: str, device: int = 0, rank: int = 0, world: int = 1, upload: bool = True,
                 chunk_rows: int = 1 << 18):
        t_path, i_path, n_path, m_path = _paths(stem)
        meta = json.load(open(m_path))
        if meta.get("format") != FORMAT:
            raise ValueError(f"{m_path}: not a {FORMAT} gallery")
        self.templates = np.load(t_path, mmap_mode="r")
        self.ids = np.load(i_path, mmap_mode="r")
        self.names = np.load(n_path, mmap_mode="r")
        self.num_students = int(meta["num_students"])
        if self.templates.shape != (self.num_students, 512) or self.templates.dtype != np.float32:
            raise ValueError(f"{t_path}: expected float32 [{self.num_students}, 512], found {self.templates.dtype} {self.templates.shape}")
        base, extra = divmod(self.num_students, world)          # same contiguous balanced shards as dist.shard_bounds
        self.lo = rank * base + min(rank, extra)
        self.hi = self.lo + base + (1 if rank < extra else 0)
        self._device, self._ctx, self._chunk, self._gen = device, None, chunk_rows, None
        if upload:
            self.upload()

    def student(self, row: int) -> Tuple[str, str]:
        return self.ids[row].decode("utf-8"), self.names[row].decode("utf-8")

    def upload(self):
        """Rows [lo, hi) -> HBM as this context's resident gallery (global ids = row numbers)."""
        import torch

        from . import _native
        self._ctx = _native.default_context(self._device)
        n = self.hi - self.lo
        dev = torch.device("cuda", self._device)
        buf = torch.empty((max(n, 1), 512), dtype=torch.float32, device=dev)
        stage = torch.empty((min(self._chunk, max(n, 1)), 512), dtype=torch.float32).pin_memory()
        for r0 in range(0, n, self._chunk):
            r1 = min(n, r0 + self._chunk)
            stage[: r1 - r0].numpy()[...] = self.templates[self.lo + r0:self.lo + r1]     # page-in from the map
            buf[r0:r1].copy_(stage[: r1 - r0], non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()                                    # stage is reused
        with self._ctx.lock:
            self._ctx.frb_gallery_upload(buf.data_ptr(), n, self.lo, 1)
            self._gen = self._ctx.gallery_generation()

    def search_batch(self, query_embeddings: np.ndarray, top_k: int = 5, threshold: float = 0.0):
        """Same contract as GalleryManager.search_batch: ([[(student_id, name, score)]], accept[P]) over this shard."""
        from . import _native
        q = np.ascontiguousarray(query_embeddings, dtype=np.float32).reshape(-1, 512)
        P, k = len(q), int(top_k)
        if P == 0 or self.hi == self.lo:
            return [[] for _ in range(P)], np.zeros(P, dtype=bool)
        scores = np.empty((P, k), np.float32)
        idx = np.empty((P, k), np.int64)
        acc = np.empty((P,), np.uint8)
        ctx = self._ctx or _native.default_context(self._device)
        with ctx.lock:
            if self._ctx is None or self._gen != ctx.gallery_generation():   # something else took the context's gallery
                self.upload()
            ctx.frb_match_host(q.ctypes.data, P, k, float(threshold), 1, scores.ctypes.data, idx.ctypes.data, acc.ctypes.data)
        out: List[List[Tuple[str, str, float]]] = []
        for p in range(P):
            row = []
            for j in range(k):
                gi = int(idx[p, j])
                if gi < 0:
                    break
                row.append((*self.student(gi), float(scores[p, j])))
            out.append(row)
        return out, acc.astype(bool)

    def search(self, query_embedding: np.ndarray, top_k: int = 5):
        return self.search_batch(np.asarray(query_embedding).reshape(1, -1), top_k)[0][0]
