"""Memory-mappable gallery layout: file contract on CPU, device search on the B200."""
import json
import os

import numpy as np
import pytest

from facerecognitionpipeline_b200 import gallery_matrix as gmx
from facerecognitionpipeline_b200.gallery_manager import GalleryManager


def _gallery(tmp_path, n=37, seed=0):
    rng = np.random.default_rng(seed)
    gm = GalleryManager(gallery_path=str(tmp_path / "g" / "students.pkl"))
    for i in range(n):
        e = rng.standard_normal((1 + i % 3, 512)).astype(np.float32)
        e /= np.linalg.norm(e, axis=1, keepdims=True)
        gm.add_student(f"STU{i:04d}", f"Student Näme {i}", e)
    return gm


def test_layout_round_trip_and_sharding(tmp_path):
    gm = _gallery(tmp_path)
    stem = str(tmp_path / "flat" / "students")
    gmx.export_gallery(gm, stem)
    assert sorted(os.listdir(tmp_path / "flat")) == ["students.ids.npy", "students.meta.json", "students.names.npy",
                                                      "students.templates.npy"]
    meta = json.load(open(stem + ".meta.json"))
    assert meta["format"] == gmx.FORMAT and meta["num_students"] == 37 and meta["dim"] == 512
    mat, ids = gm.get_gallery_embeddings()
    view = gmx.MatrixGallery(stem, upload=False)
    assert isinstance(view.templates, np.memmap)                          # mapped, not read
    assert np.array_equal(np.asarray(view.templates), mat.astype(np.float32))
    assert [view.student(r) for r in range(37)] == [(sid, gm.students[sid].name) for sid in ids]
    bounds = [(gmx.MatrixGallery(stem, rank=r, world=4, upload=False).lo, gmx.MatrixGallery(stem, rank=r, world=4, upload=False).hi)
              for r in range(4)]
    assert bounds == [(0, 10), (10, 19), (19, 28), (28, 37)]              # == dist.shard_bounds
    from facerecognitionpipeline_b200.dist import shard_bounds
    assert bounds == [shard_bounds(37, 4, r) for r in range(4)]
    with pytest.raises(ValueError):
        gmx.write_matrix(str(tmp_path / "bad"), np.zeros((2, 512), np.float32), ["a", "b"], ["only-one"])
    json.dump({"format": "other"}, open(stem + ".meta.json", "w"))
    with pytest.raises(ValueError):
        gmx.MatrixGallery(stem, upload=False)


@pytest.mark.gpu
def test_device_search_equals_gallery_manager(tmp_path):
    gm = _gallery(tmp_path, n=300, seed=1)
    stem = str(tmp_path / "flat" / "students")
    gmx.export_gallery(gm, stem)
    rng = np.random.default_rng(2)
    mat, ids = gm.get_gallery_embeddings()
    q = (mat[rng.integers(0, 300, 40)] + 0.05 * rng.standard_normal((40, 512))).astype(np.float32)
    want, want_acc = gm.search_batch(q, top_k=5, threshold=0.4)
    full = gmx.MatrixGallery(stem, chunk_rows=64)
    got, got_acc = full.search_batch(q, top_k=5, threshold=0.4)
    assert got == want and np.array_equal(got_acc, want_acc)
    assert full.search(q[0], top_k=3) == want[0][:3]
    # identity shards: each shard's hits carry global rows; their union re-ranked reproduces the full answer
    per = [gmx.MatrixGallery(stem, rank=r, world=3) for r in range(3)]
    merged = []
    for p in range(len(q)):
        cands = [t for sh in per for t in sh.search_batch(q[p:p + 1], top_k=5)[0][0]]
        cands.sort(key=lambda t: (-t[2], t[0]))
        merged.append(cands[:5])
    assert [[t[0] for t in row] for row in merged] == [[t[0] for t in row] for row in want]
