"""CPU oracle of the gallery arithmetic — TEST INFRASTRUCTURE ONLY (see oracle/backbone.py header).

PARITY STATUS: pinned.  `aggregate` reproduces the shipped gallery backups
(/root/reference/gallery/backups/*.json: 8x512 embeddings -> stored template) to <= 6e-8 and is
checked against the reference's own GalleryManager imported in the build container
(tests/golden/make_golden.py -> tests/golden/*.npz).

  search      : reference gallery_manager.py:189-205  q/(||q||+1e-8); G.q; argsort()[::-1][:k].
                The reference's order on exact ties is unspecified (reversed unstable sort); the
                oracle fixes the canonical order (score desc, row index asc) and computes scores in
                float64 from the float32 inputs so the order is well defined.
  aggregate   : reference gallery_manager.py:297-317 with the quality filter :104-122
                (gram, zeroed diagonal, row mean over n (sic), keep >= 0.70, fall back to best two).
"""
import numpy as np


def normalize_query(q: np.ndarray) -> np.ndarray:
    return q / (np.linalg.norm(q) + 1e-8)


def search(gallery: np.ndarray, query: np.ndarray, top_k: int = 5, normalize: bool = True):
    """Returns (indices [k'], scores float64 [k']) with k' = min(top_k, N)."""
    if len(gallery) == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.float64)
    q = normalize_query(query) if normalize else query
    s = (gallery.astype(np.float64) * q.astype(np.float64)[None, :]).sum(axis=1)
    order = np.lexsort((np.arange(len(s)), -s))[:top_k]
    return order.astype(np.int64), s[order]


def search_batch(gallery: np.ndarray, queries: np.ndarray, top_k: int = 5, normalize: bool = True):
    """Batched `search`; returns idx [P,k] (-1 padded), scores float64 [P,k] (-inf padded)."""
    P, N = len(queries), len(gallery)
    idx = np.full((P, top_k), -1, np.int64)
    sc = np.full((P, top_k), -np.inf, np.float64)
    if N == 0:
        return idx, sc
    q = queries.astype(np.float32)
    if normalize:
        q = np.stack([normalize_query(r) for r in q]).astype(np.float32)
    G = gallery.astype(np.float64)
    for s0 in range(0, P, 64):
        S = q[s0:s0 + 64].astype(np.float64) @ G.T
        for r in range(S.shape[0]):
            s = S[r]
            kk = min(top_k, N)
            # candidates by partition, then canonical order (score desc, index asc)
            cand = np.argpartition(-s, min(kk + 8, N - 1))[:min(kk + 9, N)] if N > kk + 16 else np.arange(N)
            thr = np.sort(s[cand])[::-1][kk - 1]
            # BLAS may round the dot product of IDENTICAL gallery rows differently depending on where the row sits in
            # its blocking (seen at N = 7: 1 ulp between two copies of one row), which would break an exact tie the
            # wrong way.  The candidates near the cut are therefore re-scored row by row with numpy's own pairwise
            # sum (the same arithmetic `search` uses: position independent), and ordered on those scores.
            cand = np.nonzero(s >= thr - 1e-9)[0]
            exact = (G[cand] * q[s0 + r].astype(np.float64)[None, :]).sum(axis=1)
            pick = np.lexsort((cand, -exact))[:kk]
            idx[s0 + r, :kk] = cand[pick]
            sc[s0 + r, :kk] = exact[pick]
    return idx, sc


def quality_filter(e: np.ndarray, min_similarity: float = 0.70) -> np.ndarray:
    if len(e) <= 2:
        return e
    g = np.dot(e, e.T)
    np.fill_diagonal(g, 0)
    avg = np.mean(g, axis=1)
    out = e[avg >= min_similarity]
    if len(out) < 2:
        out = e[np.argsort(avg)[-2:]]
    return out


def aggregate(e: np.ndarray, method: str = "mean", use_filter: bool = True) -> np.ndarray:
    if len(e) == 1:
        return e[0]
    if use_filter:
        e = quality_filter(e)
    if method == "median":
        v = np.median(e, axis=0)
    elif method == "weighted_mean":
        w = np.mean(np.dot(e, e.T), axis=1)
        w = w / np.sum(w)
        v = np.sum(e * w[:, np.newaxis], axis=0)
    else:
        v = np.mean(e, axis=0)
    return v / (np.linalg.norm(v) + 1e-8)
