#!/usr/bin/env python
"""Golden fixtures for the enrollment / embedding-generator flows, produced by the REFERENCE's own
functions (run in the build container only; the GPU box never sees /root/reference).

The reference modules cannot be imported (`insightface` / `net` are missing), so the
functions that matter are pulled out of the source files with `ast` and executed unmodified:

  flows_cases.npz
    aug/in_<i>, aug/out_<i>   enroll_students.augment_face_for_enrollment (enroll_students.py:20-48):
                              seeded RGB crops -> the 8 augmented crops the flow embeds
    names/in, names/out       EmbeddingGenerator.extract_name_from_filename (embedding_generator.py:97-106)
"""
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
    ns = dict(cv2=cv2, np=np, List=List, Dict=Dict, Tuple=Tuple, Path=Path)
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
    np.savez_compressed(os.path.join(OUT, "flows_cases.npz"), **pack)
    print({k: v.shape for k, v in pack.items()})


if __name__ == "__main__":
    main()
