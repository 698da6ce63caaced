"""Per-identity matching over a gallery of SAMPLES, device resident.

Mirror of the evaluation notebook's matcher (evaluate_models_v2.ipynb cells 3-5, called by `temp.py:19-53`):
`compute_all_similarities` scores the probe against every enrolled sample of every identity,
`aggregate_max / aggregate_mean / aggregate_topk` reduce them per identity and `identify_probe` ranks the
identities (stable sort, score descending) and applies `best_score < threshold -> None`.

The gallery dict has the notebook's shape, `{name: {"embeddings": [n_i, 512]}}`; it is uploaded once
(`frb_gallery_upload_samples`) and every query is one device call: `frb_match_identities` (tensor-core filter over
the samples -> exact f64 aggregates of the candidate identities -> proof -> exact scan where the proof fails) or
`frb_identity_scores` (the full [P][S] score matrix `identify_probe` returns as a dict).

cosine_similarity (cell 3) takes the plain dot product when both norms are within 0.01 of 1 and divides by
n1*n2 otherwise.  FaceEmbedder output is unit-norm to f32 rounding, which is the regime reproduced to 1e-6:
probes whose norm is within 0.01 of 1 are scored as they are, any other probe is divided by its norm first.
Gallery samples must be unit-norm to within 1e-3 (ValueError otherwise): the per-pair 1/n2 of a sample that is
"nearly" unit is not applied.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np

AGGREGATIONS = {"max": 0, "mean": 1, "topk": 2}
MAX_SAMPLES = 64


class IdentityGallery:
    """Device-resident sample gallery; replaces the context's template gallery while it is in use."""

    def __init__(self, gallery: Dict[str, Dict], device: int = 0):
        self.names: List[str] = list(gallery)
        rows, seg = [], [0]
        for name in self.names:
            e = np.asarray(gallery[name]["embeddings"], dtype=np.float32).reshape(-1, 512)
            if len(e) > MAX_SAMPLES:
                raise ValueError(f"identity {name!r} has {len(e)} samples; at most {MAX_SAMPLES} are supported")
            rows.append(e)
            seg.append(seg[-1] + len(e))
        self.samples = np.ascontiguousarray(np.concatenate(rows) if rows else np.zeros((0, 512), np.float32))
        self.seg = np.asarray(seg, dtype=np.int64)
        if len(self.samples):
            n = np.linalg.norm(self.samples.astype(np.float64), axis=1)
            if np.abs(n - 1.0).max() > 1e-3:
                raise ValueError("gallery samples must be L2-normalised embeddings (FaceEmbedder output)")
        self._device, self._ctx, self._gen = device, None, None

    # ---- device plumbing
    def _context(self):
        from . import _native
        if self._ctx is None:
            self._ctx = _native.default_context(self._device)
        return self._ctx

    def _resident(self):
        """Upload the samples unless this object's upload is still the context's resident gallery (the ctx counts
        uploads: `frb_gallery_generation`; object identity is NOT used, ids are recycled).  Lock held by the caller."""
        ctx = self._context()
        if self._gen is None or self._gen != ctx.gallery_generation():
            ctx.frb_gallery_upload_samples(self.samples.ctypes.data, len(self.samples), self.seg.ctypes.data,
                                           len(self.names), 0)
            self._gen = ctx.gallery_generation()
        return ctx

    @staticmethod
    def _prepare(probes: np.ndarray) -> np.ndarray:
        q = np.array(probes, dtype=np.float32).reshape(-1, 512)
        n = np.linalg.norm(q.astype(np.float64), axis=1)
        off = np.abs(n - 1.0) >= 0.01                                   # cell 3: the dot / (n1 * n2) branch
        if off.any():
            q[off] = (q[off].astype(np.float64) / n[off, None]).astype(np.float32)
        return np.ascontiguousarray(q)

    @staticmethod
    def _agg(aggregation: str) -> int:
        return AGGREGATIONS.get(aggregation, 0)                          # identify_probe: anything else -> aggregate_max

    # ---- queries
    def identity_scores(self, probes: np.ndarray, aggregation: str = "mean", k: int = 3) -> np.ndarray:
        """[P][S] f32: row p is `identify_probe(...)[2]` in gallery order (an identity without samples scores -1)."""
        import torch
        q = self._prepare(probes)
        P, S = len(q), len(self.names)
        if P == 0 or S == 0:
            return np.zeros((P, S), np.float32)
        dev = torch.device("cuda", self._device)
        d_q = torch.from_numpy(q).to(dev)
        out = torch.empty((P, S), dtype=torch.float32, device=dev)
        with self._context().lock:
            ctx = self._resident()
            ctx.frb_identity_scores(d_q.data_ptr(), P, 0, self._agg(aggregation), max(int(k), 1), out.data_ptr(), S,
                                    torch.cuda.current_stream(dev).cuda_stream)
            torch.cuda.current_stream(dev).synchronize()
        return out.cpu().numpy()

    def rank_batch(self, probes: np.ndarray, top_k: int = 1, threshold: float = 0.0, aggregation: str = "mean",
                   k: int = 3) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """(identity index [P][top_k] i64 (-1 = none), score [P][top_k] f32, accept [P] bool)."""
        import torch
        q = self._prepare(probes)
        P, S, K = len(q), len(self.names), int(top_k)
        idx = np.full((P, K), -1, np.int64)
        sc = np.full((P, K), -np.inf, np.float32)
        acc = np.zeros((P,), bool)
        if P == 0 or S == 0:
            return idx, sc, acc
        dev = torch.device("cuda", self._device)
        d_q = torch.from_numpy(q).to(dev)
        d_sc = torch.empty((P, K), dtype=torch.float32, device=dev)
        d_ix = torch.empty((P, K), dtype=torch.int64, device=dev)
        d_ac = torch.empty((P,), dtype=torch.uint8, device=dev)
        with self._context().lock:
            ctx = self._resident()
            ctx.frb_match_identities(d_q.data_ptr(), P, K, float(threshold), 0, self._agg(aggregation), max(int(k), 1),
                                     d_sc.data_ptr(), d_ix.data_ptr(), d_ac.data_ptr(),
                                     torch.cuda.current_stream(dev).cuda_stream)
            torch.cuda.current_stream(dev).synchronize()
        return d_ix.cpu().numpy(), d_sc.cpu().numpy(), d_ac.cpu().numpy().astype(bool)

    def identify_batch(self, probes: np.ndarray, threshold: float, aggregation: str = "mean",
                       k: int = 3) -> List[Tuple[Optional[str], float]]:
        """[(predicted identity or None, best score)] - `identify_probe(...)[:2]` for every probe."""
        if not self.names:
            return [(None, -1)] * len(np.asarray(probes).reshape(-1, 512))
        idx, sc, acc = self.rank_batch(probes, 1, threshold, aggregation, k)
        return [(self.names[int(i)] if a else None, float(s)) for i, s, a in zip(idx[:, 0], sc[:, 0], acc)]


def identify_probe(probe_embedding: np.ndarray, gallery, threshold: float, aggregation: str = "mean",
                   k: int = 3) -> Tuple[Optional[str], float, Dict[str, float]]:
    """Drop-in for the notebook's `identify_probe` (cell 5).  `gallery` is the notebook's dict or an
    `IdentityGallery` built from it (build it once when many probes are identified)."""
    g = gallery if isinstance(gallery, IdentityGallery) else IdentityGallery(gallery)
    if not g.names:
        return None, -1, {}
    row = g.identity_scores(np.asarray(probe_embedding).reshape(1, 512), aggregation, k)[0]
    scores = {n: float(s) for n, s in zip(g.names, row)}
    (name, best), = g.identify_batch(np.asarray(probe_embedding).reshape(1, 512), threshold, aggregation, k)
    return name, best, scores
