"""Pins oracle/gallery.py and the product's host-side template code against the reference's own
outputs (golden fixtures produced by tests/golden/make_golden.py from /root/reference)."""
import json
import os

import numpy as np
import pytest

from oracle import gallery as og
from facerecognitionpipeline_b200 import templates


@pytest.fixture(scope="module")
def backups(golden_dir):
    z = np.load(os.path.join(golden_dir, "gallery_backups.npz"))
    tags = sorted({k.split("/")[0] for k in z.files})
    return z, tags


def test_backups_shape(backups):
    z, tags = backups
    assert len(tags) == 5
    for t in tags:
        assert z[t + "/emb"].shape == (23, 8, 512) and z[t + "/tpl"].shape == (23, 512)
        n = np.linalg.norm(z[t + "/emb"], axis=2)
        assert np.abs(n - 1).max() < 1e-5   # stored embeddings are unit norm


@pytest.mark.parametrize("impl", ["oracle", "product"])
def test_mean_template_reproduces_shipped_backups(backups, impl):
    """filter + mean + renorm reproduces every stored template to <= 1e-7 (1 f32 ulp-ish)."""
    z, tags = backups
    worst = 0.0
    for t in tags:
        for e, tpl in zip(z[t + "/emb"], z[t + "/tpl"]):
            got = og.aggregate(e, "mean") if impl == "oracle" else templates.gallery_template(e, "mean", verbose=False)
            worst = max(worst, float(np.abs(got - tpl).max()))
    assert worst <= 1e-7, worst


@pytest.mark.parametrize("impl", ["oracle", "product"])
def test_aggregate_matches_reference_outputs(golden_dir, impl):
    """Every method incl. weighted_mean / median / unknown-method fallback, the quality filter and the
    outlier filter, against outputs of the reference's own GalleryManager."""
    z = np.load(os.path.join(golden_dir, "aggregate_cases.npz"))
    for c in range(int(z["num_cases"])):
        method = str(z[f"c{c}/method"])
        e = z[f"c{c}/in"]
        if impl == "oracle":
            out, filt = og.aggregate(e.copy(), method), og.quality_filter(e.copy())
        else:
            out = templates.gallery_template(e.copy(), method, verbose=False)
            filt = templates.quality_filter(e.copy(), verbose=False)
            np.testing.assert_array_equal(templates.drop_outliers(e.copy()), z[f"c{c}/outliers"])
        np.testing.assert_array_equal(filt, z[f"c{c}/filtered"])
        np.testing.assert_allclose(out, z[f"c{c}/out"], rtol=0, atol=1e-7)


def test_embedder_template_errors():
    with pytest.raises(ValueError):
        templates.embedder_template(np.zeros((0, 512)))
    with pytest.raises(ValueError):
        templates.embedder_template(np.ones((3, 512), np.float32), "bogus")
    one = np.arange(512, dtype=np.float32)[None]
    assert templates.embedder_template(one) is not None and np.array_equal(templates.embedder_template(one), one[0])


def test_oracle_search_matches_reference_search(golden_dir, backups):
    z, _ = backups
    s = np.load(os.path.join(golden_dir, "search_cases.npz"))
    tag = str(s["gallery_tag"])
    G, ids = z[tag + "/tpl"], z[tag + "/ids"]
    for q, exp_ids, exp_sc in zip(s["probes"], s["ids"], s["scores"]):
        idx, sc = og.search(G, q, 5)
        assert [str(ids[i]) for i in idx] == [str(x) for x in exp_ids]
        np.testing.assert_allclose(sc, exp_sc, atol=2e-6)
    bi, bs = og.search_batch(G, s["probes"], 5)
    for p, q in enumerate(s["probes"]):
        idx, sc = og.search(G, q, 5)
        assert np.array_equal(bi[p], idx) and np.allclose(bs[p], sc, atol=1e-12)


def test_verify_enrollment_property(backups):
    """enroll_students.verify_enrollment (:365-373): a student's own embedding retrieves that student at rank 1."""
    z, tags = backups
    for t in tags:
        G = z[t + "/tpl"]
        for sidx, e in enumerate(z[t + "/emb"]):
            idx, _ = og.search(G, e[0], 1)
            assert idx[0] == sidx


def test_search_edge_cases():
    idx, sc = og.search(np.zeros((0, 512), np.float32), np.ones(512, np.float32), 5)
    assert len(idx) == 0
    G = np.eye(512, dtype=np.float32)[:3]
    G = np.vstack([G, G[1:2]])          # duplicate row -> exact tie, lower index first
    idx, sc = og.search(G, G[1], 5)
    assert idx.tolist()[:2] == [1, 3] and len(idx) == 4


def test_consensus_reproduces_recorded_track(golden_dir):
    """_aggregate_matches on the recorded per-frame matches gives the recorded decision and the recorded
    confidence exactly (mean of the 12 frame scores)."""
    from facerecognitionpipeline_b200.face_matcher import consensus, best_candidate
    rec = json.load(open(os.path.join(golden_dir, "track_001.json")))
    out = consensus(rec["frame_matches"], 0.5)
    assert out is not None and rec["recognized"]
    assert out["student_id"] == rec["student_id"] and out["name"] == rec["name"]
    assert out["confidence"] == rec["confidence"]
    assert out["total_frames_evaluated"] == rec["num_frames"]
    assert best_candidate(rec["frame_matches"])["student_id"] == rec["student_id"]
    # decision rules
    fm = rec["frame_matches"]
    assert consensus(fm[:2], 0.5) is None                      # fewer than 3 quality frames
    assert consensus(fm, 0.99) is None                          # mean below threshold
    split = [dict(m, student_id=("A" if i % 2 else "B")) for i, m in enumerate(fm[:4])]
    assert consensus(split, 0.1) is None                        # 50/50 is not a majority
