#!/usr/bin/env python
"""Golden fixtures for the enrollment / embedding-generator flows, produced by the REFERENCE's own
functions (run in the build container only; the GPU box never sees /root/reference).

The reference modules cannot be imported (`insightface` / `net` are missing), so the
functions that matter are pulled out of the source files with `ast` and executed unmodified:

  flows_cases.npz
    aug/in_<i>, aug/out_<i>   enroll_students.augment_face_for_enrollment (enroll_students.py:20-48):
                              seeded RGB crops -> the 8 augmented crops the flow embeds
    names/in, names/out       EmbeddingGenerator.extract_name_from_filename (embedding_generator.py:97-106)
    tracks/*                  FaceMatcher._aggregate_matches / _get_best_candidate (face_matcher.py:321-385) executed on
                              seeded synthetic tracks: per-frame top-1 (identity, f32 score) -> the decision fields
    identify/*                evaluate_models_v2.ipynb cells 3-5 (cosine_similarity, aggregate_*, identify_probe) executed
                              on a seeded ragged sample gallery: per probe and aggregation the identity score vector,
                              the predicted identity index (-1 = rejected) and the best score
"""
import json
import ast
import os
import types

import cv2
import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def extract(path, name, cls=None):
    tree = ast.parse(open(path).read())
    body = tree.body
    if cls:
        body = next(n for n in body if isinstance(n, ast.ClassDef) and n.name == cls).body
    fn = next(n for n in body if isinstance(n, ast.FunctionDef) and n.name == name)
    mod = ast.Module(body=[fn], type_ignores=[])
    from pathlib import Path
    from typing import Dict, List, Tuple
    from collections import Counter
    from typing import Optional
    ns = dict(cv2=cv2, np=np, List=List, Dict=Dict, Tuple=Tuple, Path=Path, Counter=Counter, Optional=Optional)
    exec(compile(mod, path, "exec"), ns)
    return ns[name]


def main():
    augment = extract(f"{REF}/enroll_students.py", "augment_face_for_enrollment")
    pack = {}
    rng = np.random.default_rng(77)
    for i, S in enumerate((112, 64, 96)):
        img = cv2.GaussianBlur(rng.integers(0, 256, (S, S, 3), dtype=np.uint8), (0, 0), 1.5)
        np.random.seed(0)  # the (unused, beyond index 8) noise variant draws from the global RNG
        out = augment(img, num_augmentations=8)
        pack[f"aug/in_{i}"] = img
        pack[f"aug/out_{i}"] = np.stack(out)
    name_fn = extract(f"{REF}/embedding_generator.py", "extract_name_from_filename", cls="EmbeddingGenerator")
    names = ["alice_smith_001_f3.jpg", "bob_12.png", "007_bond.jpg", "carol.jpeg", "dave_lee_x_9_9.jpg", "lfw_Aaron_Eckhart_0001.jpg"]
    pack["names/in"] = np.array(names)
    pack["names/out"] = np.array([name_fn(types.SimpleNamespace(), n) for n in names])
    # ---- the notebook's own matcher (cells 3, 4, 5 are pure function definitions)
    nb = json.load(open(f"{REF}/evaluate_models_v2.ipynb"))
    from typing import Dict, List, Optional, Tuple
    ns = dict(np=np, Dict=Dict, List=List, Optional=Optional, Tuple=Tuple)
    for ci in (3, 4, 5):
        exec("".join(nb["cells"][ci]["source"]), ns)
    rng = np.random.default_rng(99)
    counts = [1, 8, 3, 5, 8, 2, 8, 40, 4, 8, 6, 1]
    centres = rng.standard_normal((len(counts), 512))
    gallery, rows = {}, []
    for i, n in enumerate(counts):
        e = centres[i][None] + 0.9 * rng.standard_normal((n, 512))
        e = (e / np.linalg.norm(e, axis=1, keepdims=True)).astype(np.float32)
        gallery[f"person_{i:02d}"] = {"embeddings": e}
        rows.append(e)
    probes = []
    for i in range(len(counts)):
        p = centres[i] + 1.1 * rng.standard_normal(512)
        probes.append(p / np.linalg.norm(p))
    probes += [rng.standard_normal(512) for _ in range(4)]                     # impostors
    probes[-1] = probes[-1] / np.linalg.norm(probes[-1])
    probes = np.array(probes, dtype=np.float32)
    probes[1] *= 1.7                                                            # not unit norm: exercises the cosine branch
    pack["identify/samples"] = np.concatenate(rows)
    pack["identify/counts"] = np.array(counts)
    pack["identify/probes"] = probes
    names = list(gallery)
    for agg in ("max", "mean", "topk", "bogus"):
        S, pred, best = [], [], []
        for p in probes:
            name, score, scores = ns["identify_probe"](p, gallery, threshold=0.3, aggregation=agg, k=3)
            S.append([scores[n] for n in names])
            pred.append(-1 if name is None else names.index(name))
            best.append(score)
        pack[f"identify/{agg}/scores"] = np.array(S, dtype=np.float64)
        pack[f"identify/{agg}/pred"] = np.array(pred)
        pack[f"identify/{agg}/best"] = np.array(best, dtype=np.float64)
    # ---- track-level consensus: the reference's own methods on synthetic tracks
    agg_fn = extract(f"{REF}/face_matcher.py", "_aggregate_matches", cls="FaceMatcher")
    cand_fn = extract(f"{REF}/face_matcher.py", "_get_best_candidate", cls="FaceMatcher")
    rng = np.random.default_rng(2024)
    seg, ids, scores, thr_list, rows = [0], [], [], [], []
    for t in range(400):
        F = int(rng.choice([0, 1, 2, 3, 4, 5, 7, 8, 9, 12, 16, 17, 30, 64, 129, 150, 300]))
        pool = rng.choice(50, size=int(rng.integers(1, 5)), replace=False)
        fid = rng.choice(pool, size=F, p=None)
        lo = float(rng.choice([0.3, 0.5, 0.56]))
        sc = rng.uniform(lo, 0.95, size=F).astype(np.float32)
        if F and rng.random() < 0.3:
            fid[rng.integers(0, F, max(1, F // 5))] = -1            # frames without any match (skipped)
        thr = float(rng.choice([0.4, 0.5, 0.6, 0.7]))
        fm = [dict(student_id=f"STU{i:04d}", name=f"n{i}", score=float(s)) for i, s in zip(fid, sc) if i >= 0]
        self_ = types.SimpleNamespace(similarity_threshold=thr)
        a = agg_fn(self_, fm, {}) if fm else None
        c = cand_fn(self_, fm, {}) if fm else None
        rows.append([
            1 if a else 0, int(a["student_id"][3:]) if a else -1, a["confidence"] if a else 0.0,
            a["consensus_strength"] if a else 0.0, a["num_quality_frames"] if a else 0, len(fm),
            int(c["student_id"][3:]) if c else -1, c["confidence"] if c else 0.0, c["num_quality_frames"] if c else 0])
        ids += list(fid); scores += list(sc); seg.append(seg[-1] + F); thr_list.append(thr)
    pack["tracks/seg"] = np.array(seg, np.int64)
    pack["tracks/ids"] = np.array(ids, np.int64)
    pack["tracks/scores"] = np.array(scores, np.float32)
    pack["tracks/thr"] = np.array(thr_list, np.float64)
    pack["tracks/out"] = np.array(rows, np.float64)    # recognized, winner, confidence, strength, nq, total, cand, cand_conf, cand_nq
    np.savez_compressed(os.path.join(OUT, "flows_cases.npz"), **pack)
    print({k: v.shape for k, v in pack.items()})


if __name__ == "__main__":
    main()
