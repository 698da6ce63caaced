"""Device enrollment aggregation (`frb_aggregate_templates`) against the reference's own
GalleryManager outputs (tests/golden/aggregate_cases.npz: all three methods + the silent fallback, n = 1..40,
including cases where the quality filter drops rows or falls back to the two best) and the 5 shipped
full-embedding backups (23 students x 8 embeddings -> stored template).  Tolerance 2e-6 absolute on unit
vectors: the reference's BLAS/pairwise fp32 summation order is not reproduced, everything else is."""
import os

import numpy as np
import pytest
import torch

from facerecognitionpipeline_b200 import templates

pytestmark = pytest.mark.gpu


def test_golden_aggregate_cases(golden_dir):
    g = np.load(os.path.join(golden_dir, "aggregate_cases.npz"))
    cases = [f"c{i}" for i in range(int(g["num_cases"]))]
    by_method = {}
    for c in cases:
        by_method.setdefault(str(g[f"{c}/method"]), []).append(c)
    assert set(by_method) == {"mean", "median", "weighted_mean", "bogus"}
    for method, cs in by_method.items():
        rows = np.concatenate([g[f"{c}/in"] for c in cs])
        counts = [len(g[f"{c}/in"]) for c in cs]
        out, kept = templates.aggregate_on_device(rows, counts, method=method)
        out, kept = out.cpu().numpy(), kept.cpu().numpy()
        for i, c in enumerate(cs):
            want = g[f"{c}/out"]
            assert np.abs(out[i] - want).max() < 2e-6, (method, c, np.abs(out[i] - want).max())
            n_in = counts[i]
            assert kept[i] == (n_in if n_in <= 2 else len(g[f"{c}/filtered"]))
            host = templates.gallery_template(g[f"{c}/in"], method, verbose=False)
            assert np.abs(out[i] - host).max() < 2e-6


def test_shipped_backups_and_resident_gallery(ctx, golden_dir):
    """Every stored template of the 5 shipped galleries is reproduced from its 8 stored embeddings, and the
    templates uploaded straight from the device retrieve their own students (enroll_students.verify_enrollment)."""
    g = np.load(os.path.join(golden_dir, "gallery_backups.npz"))
    tags = sorted({k.split("/")[0] for k in g.files})
    assert len(tags) == 5
    for tag in tags:
        emb, tpl = g[f"{tag}/emb"], g[f"{tag}/tpl"]                  # [23,8,512], [23,512]
        S, n = emb.shape[0], emb.shape[1]
        d_emb = torch.from_numpy(emb.reshape(S * n, 512)).cuda()
        out, kept = templates.aggregate_on_device(d_emb, [n] * S, method="mean", upload_as_gallery=True)
        assert np.abs(out.cpu().numpy() - tpl).max() < 2e-6
        assert (kept.cpu().numpy() == n).all()
        sc = np.empty((S, 3), np.float32); ix = np.empty((S, 3), np.int64); ac = np.empty((S,), np.uint8)
        q = np.ascontiguousarray(emb[:, 0, :])
        ctx.frb_match_host(q.ctypes.data, S, 3, 0.5, 1, sc.ctypes.data, ix.ctypes.data, ac.ctypes.data)
        assert (ix[:, 0] == np.arange(S)).all() and ac.all()


def test_edge_cases():
    rng = np.random.default_rng(0)
    base = rng.standard_normal(512)
    e = base[None] + 0.2 * rng.standard_normal((9, 512))
    e = (e / np.linalg.norm(e, axis=1, keepdims=True)).astype(np.float32)
    out, kept = templates.aggregate_on_device(np.concatenate([e[:1], e[:2], e]), [1, 2, 9, 0][:3], method="median")
    out = out.cpu().numpy()
    assert np.array_equal(out[0], e[0])                                   # single row: returned as is, not re-normalised
    assert np.abs(out[1] - templates.gallery_template(e[:2], "median", verbose=False)).max() < 2e-6
    assert np.abs(out[2] - templates.gallery_template(e, "median", verbose=False)).max() < 2e-6
    with pytest.raises(ValueError):
        templates.aggregate_on_device(e, [4, 4], method="mean")           # counts do not add up
    with pytest.raises(ValueError):
        templates.aggregate_on_device(np.zeros((65, 512), np.float32), [65])
