"""Template (per-identity) aggregation of a handful of embeddings — host side, n <= ~40 rows.

Semantics follow the reference exactly, including its quirks:
  * GalleryManager._filter_quality_embeddings (gallery_manager.py:104-122): gram matrix with a zeroed
    diagonal, row mean taken over ALL n columns (so it divides by n, not n-1), keep rows whose mean is
    >= 0.70, fall back to the two best rows when fewer than two survive; skipped for n <= 2.
  * GalleryManager._aggregate_embeddings (gallery_manager.py:297-317) / FaceEmbedder.aggregate_embeddings
    (face_embedder.py:202-225): a single row is returned as is (not re-normalised); otherwise
    mean / median / weighted_mean (weights = row means of the gram matrix INCLUDING the diagonal,
    normalised to sum 1) followed by x / (||x|| + 1e-8).
SURVEY §8(f) ranks a device version of this as the next row after the hot path; at n <= 40 it is
not roofline-relevant, so it stays in numpy here.
"""
from __future__ import annotations

import numpy as np

QUALITY_MIN_SIMILARITY = 0.70


def quality_filter(embeddings: np.ndarray, min_similarity: float = QUALITY_MIN_SIMILARITY, verbose: bool = True) -> np.ndarray:
    n = len(embeddings)
    if n <= 2:
        return embeddings
    gram = embeddings @ embeddings.T
    np.fill_diagonal(gram, 0)
    row_score = gram.mean(axis=1)          # divides by n on purpose (reference behaviour)
    keep = row_score >= min_similarity
    kept = embeddings[keep]
    if len(kept) < 2:
        kept = embeddings[np.argsort(row_score)[-2:]]
    if verbose:
        print(f"    Quality filter: kept {len(kept)}/{n} embeddings (threshold={min_similarity})")
    return kept


def reduce_rows(embeddings: np.ndarray, method: str, strict: bool) -> np.ndarray:
    if method == "mean":
        v = embeddings.mean(axis=0)
    elif method == "median":
        v = np.median(embeddings, axis=0)
    elif method == "weighted_mean":
        w = (embeddings @ embeddings.T).mean(axis=1)
        w = w / w.sum()
        v = (embeddings * w[:, None]).sum(axis=0)
    elif strict:
        raise ValueError(f"Unknown aggregation method: {method}")
    else:                                   # GalleryManager silently falls back to mean
        v = embeddings.mean(axis=0)
    return v / (np.linalg.norm(v) + 1e-8)


def gallery_template(embeddings: np.ndarray, method: str = "mean", verbose: bool = True) -> np.ndarray:
    """GalleryManager._aggregate_embeddings."""
    if len(embeddings) == 1:
        return embeddings[0]
    return reduce_rows(quality_filter(embeddings, verbose=verbose), method, strict=False)


def embedder_template(embeddings: np.ndarray, method: str = "mean") -> np.ndarray:
    """FaceEmbedder.aggregate_embeddings (no quality filter, unknown method is an error)."""
    if len(embeddings) == 0:
        raise ValueError("Cannot aggregate empty embeddings")
    if len(embeddings) == 1:
        return embeddings[0]
    return reduce_rows(embeddings, method, strict=True)


def drop_outliers(embeddings: np.ndarray, threshold: float = 0.7) -> np.ndarray:
    """GalleryManager._remove_outliers (gallery_manager.py:319-330): keep rows whose mean gram-row
    similarity (diagonal included) is at least threshold x the median of those means."""
    if len(embeddings) <= 2:
        return embeddings
    row_score = (embeddings @ embeddings.T).mean(axis=1)
    return embeddings[row_score >= np.median(row_score) * threshold]


# ---------------------------------------------------------------------------------------------------
# Device version for bulk gallery builds: embeddings stay in HBM between frb_embed,
# aggregation and frb_gallery_upload.  GalleryManager.add_student keeps using the numpy functions above
# (they are the reference's arithmetic to the bit); this path agrees with them to ~1e-6.
METHOD_CODES = {"mean": 0, "median": 1, "weighted_mean": 2}
DEVICE_MAX_ROWS = 64


def aggregate_on_device(embeddings, counts, method: str = "mean", min_similarity: float = QUALITY_MIN_SIMILARITY,
                        device: int = 0, upload_as_gallery: bool = False, first_global_id: int = 0):
    """Templates of S identities in one kernel launch.

    embeddings: [T, 512] float32, a CUDA torch tensor (used in place) or a numpy array (copied once);
    counts: rows per identity, in order (sum == T, each <= 64).  Unknown methods fall back to mean like
    GalleryManager does.  Returns (templates [S,512] CUDA tensor, kept [S] CUDA int32 tensor); with
    upload_as_gallery the templates also become the context's resident gallery without leaving the device."""
    import ctypes as C

    import torch

    from . import _native
    ctx = _native.default_context(device)
    dev = torch.device("cuda", device)
    emb = embeddings if isinstance(embeddings, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(embeddings, dtype=np.float32))
    emb = emb.to(dev, dtype=torch.float32).contiguous().reshape(-1, 512)
    counts = np.asarray(counts, dtype=np.int64)
    if counts.sum() != emb.shape[0]:
        raise ValueError("counts do not add up to the number of embedding rows")
    if len(counts) and counts.max() > DEVICE_MAX_ROWS:
        raise ValueError(f"at most {DEVICE_MAX_ROWS} embeddings per identity on the device path")
    S = len(counts)
    seg = torch.from_numpy(np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)).to(dev)
    out = torch.empty((S, 512), dtype=torch.float32, device=dev)
    kept = torch.empty((S,), dtype=torch.int32, device=dev)
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    ctx.frb_aggregate_templates(emb.data_ptr(), seg.data_ptr(), S, int(counts.max()) if S else 0, METHOD_CODES.get(method, 0),
                                float(min_similarity), out.data_ptr(), kept.data_ptr(), st)
    if upload_as_gallery and S:
        torch.cuda.current_stream(dev).synchronize()
        ctx.frb_gallery_upload(out.data_ptr(), S, int(first_global_id), 1)   # bumps the ctx's gallery generation
    return out, kept
