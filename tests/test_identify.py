"""Per-identity matching over gallery samples.  The golden vectors are outputs of the evaluation
notebook's own cells (tests/golden/make_golden_flows.py executes evaluate_models_v2.ipynb cells 3-5)."""
import os

import numpy as np
import pytest

from oracle import identify as oid

AGGS = ("max", "mean", "topk", "bogus")


def _golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "flows_cases.npz"))
    counts = g["identify/counts"]
    seg = np.concatenate([[0], np.cumsum(counts)])
    gallery = {f"person_{i:02d}": {"embeddings": g["identify/samples"][seg[i]:seg[i + 1]]} for i in range(len(counts))}
    return g, gallery


def test_oracle_pinned_by_notebook_outputs(golden_dir):
    g, gallery = _golden(golden_dir)
    names = list(gallery)
    for agg in AGGS:
        for p, probe in enumerate(g["identify/probes"]):
            name, best, scores = oid.identify_probe(probe, gallery, threshold=0.3, aggregation=agg, k=3)
            assert np.array_equal(np.array([scores[n] for n in names], np.float64), g[f"identify/{agg}/scores"][p])
            assert (-1 if name is None else names.index(name)) == g[f"identify/{agg}/pred"][p]
            assert best == g[f"identify/{agg}/best"][p]
    assert oid.identify_probe(g["identify/probes"][0], {}, 0.3) == (None, -1, {})
    assert oid.aggregate([], "mean") == -1


def test_host_module_contract_without_gpu():
    from facerecognitionpipeline_b200.identify import IdentityGallery, identify_probe
    assert identify_probe(np.ones(512, np.float32), {}, 0.3) == (None, -1, {})
    with pytest.raises(ValueError):
        IdentityGallery({"a": {"embeddings": np.ones((2, 512), np.float32)}})          # not unit norm
    with pytest.raises(ValueError):
        IdentityGallery({"a": {"embeddings": np.full((65, 512), 512 ** -0.5, np.float32)}})
    q = np.ones((2, 512), np.float32)
    q[0] *= 512 ** -0.5
    prep = IdentityGallery._prepare(q)
    assert np.array_equal(prep[0], q[0]) and abs(np.linalg.norm(prep[1]) - 1) < 1e-6  # only the off-norm probe is divided


@pytest.mark.gpu
def test_device_matches_notebook_golden(golden_dir):
    from facerecognitionpipeline_b200.identify import IdentityGallery, identify_probe
    g, gallery = _golden(golden_dir)
    names = list(gallery)
    ig = IdentityGallery(gallery)
    probes = g["identify/probes"]
    for agg in AGGS:
        S = ig.identity_scores(probes, agg, 3)
        assert np.abs(S - g[f"identify/{agg}/scores"]).max() < 1e-6
        got = ig.identify_batch(probes, 0.3, agg, 3)
        pred = [-1 if n is None else names.index(n) for n, _ in got]
        assert pred == g[f"identify/{agg}/pred"].tolist()
        assert np.abs(np.array([s for _, s in got]) - g[f"identify/{agg}/best"]).max() < 1e-6
    name, best, scores = identify_probe(probes[0], gallery, 0.3, "max")
    assert name == names[g["identify/max/pred"][0]] and list(scores) == names


def _big(S, seed, max_n=8, spread=0.8):
    rng = np.random.default_rng(seed)
    counts = rng.integers(0, max_n + 1, S)
    counts[:3] = [0, 1, max_n]
    centres = rng.standard_normal((S, 512))
    gallery = {}
    for i, n in enumerate(counts):
        e = centres[i][None] + spread * rng.standard_normal((n, 512))
        e = (e / np.linalg.norm(e, axis=1, keepdims=True)).astype(np.float32)
        gallery[f"id{i}"] = {"embeddings": e}
    pick = rng.choice(np.nonzero(counts)[0], 48, replace=False)
    probes = centres[pick] + 1.0 * rng.standard_normal((48, 512))
    probes = np.concatenate([probes, rng.standard_normal((16, 512))])
    probes = (probes / np.linalg.norm(probes, axis=1, keepdims=True)).astype(np.float32)
    return gallery, probes


@pytest.mark.gpu
@pytest.mark.parametrize("S", [300, 3000])          # 3000 identities -> > 4096 samples: tensor-core filter + proof path
def test_device_matches_oracle_ranking(ctx, S):
    from facerecognitionpipeline_b200.identify import IdentityGallery
    gallery, probes = _big(S, seed=S)
    names = list(gallery)
    ig = IdentityGallery(gallery)
    for agg in ("max", "mean", "topk"):
        idx, sc, acc = ig.rank_batch(probes, top_k=5, threshold=0.25, aggregation=agg, k=3)
        full = ig.identity_scores(probes, agg, 3).astype(np.float64)
        for p, probe in enumerate(probes):
            _, _, scores = oid.identify_probe(probe, gallery, 0.25, agg, 3)
            want = np.array([scores[n] for n in names], np.float64)
            assert np.abs(full[p] - want).max() < 1e-6
            # canonical ranking (score desc, identity index asc) of the device's own exact scores
            order = np.lexsort((np.arange(S), -full[p]))[:5]
            assert idx[p].tolist() == order.tolist(), (agg, p)
            assert np.abs(sc[p] - full[p][order]).max() < 1e-6
            # top-1 and accept/reject agree with the notebook wherever its f32 scores separate the two best
            srt = np.sort(want)[::-1]
            if srt[0] - srt[1] > 1e-5:
                assert idx[p, 0] == int(np.argmax(want))
            if abs(srt[0] - 0.25) > 1e-5:
                assert bool(acc[p]) == bool(srt[0] >= 0.25)
    assert ctx._lib.frb_match_last_flagged(ctx.handle) >= 0
