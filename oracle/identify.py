"""CPU restatement of the reference's per-identity matcher — TEST INFRASTRUCTURE ONLY.

Follows evaluate_models_v2.ipynb cells 3-5 (the notebook defines the helpers `temp.py:19-53` calls):
  cell 3  cosine_similarity(e1, e2): plain dot when both norms are within 0.01 of 1, else dot / (n1*n2);
          compute_all_similarities: every gallery SAMPLE of every identity
  cell 4  aggregate_max / aggregate_mean / aggregate_topk (mean of the min(k, n) largest); -1 for an empty list
  cell 5  identify_probe: identity score = aggregate over that identity's samples (unknown aggregation -> max);
          identities sorted by score descending with Python's stable sort (ties keep gallery dict order);
          best below the threshold -> (None, best_score, scores)
Pinned by tests/golden/flows_cases.npz (`identify/*`), produced by executing the notebook's own cells
(tests/golden/make_golden_flows.py).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np


def cosine_similarity(emb1: np.ndarray, emb2: np.ndarray) -> float:
    n1, n2 = np.linalg.norm(emb1), np.linalg.norm(emb2)
    if abs(n1 - 1.0) < 0.01 and abs(n2 - 1.0) < 0.01:
        return np.dot(emb1, emb2)
    return np.dot(emb1, emb2) / (n1 * n2)


def aggregate(similarities: List[float], aggregation: str, k: int = 3) -> float:
    if not len(similarities):
        return -1
    if aggregation == "mean":
        return np.mean(similarities)
    if aggregation == "topk":
        return np.mean(sorted(similarities, reverse=True)[:min(k, len(similarities))])
    return max(similarities)


def identify_probe(probe: np.ndarray, gallery: Dict[str, Dict], threshold: float, aggregation: str = "mean",
                   k: int = 3) -> Tuple[Optional[str], float, Dict[str, float]]:
    scores = {name: aggregate([cosine_similarity(probe, g) for g in data["embeddings"]], aggregation, k)
              for name, data in gallery.items()}
    if not scores:
        return None, -1, {}
    best_name, best_score = sorted(scores.items(), key=lambda kv: kv[1], reverse=True)[0]
    if best_score < threshold:
        return None, best_score, scores
    return best_name, best_score, scores


def rank_identities(scores: Dict[str, float], top_k: int) -> List[Tuple[str, float]]:
    return sorted(scores.items(), key=lambda kv: kv[1], reverse=True)[:top_k]
