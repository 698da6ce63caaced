"""Track-level consensus.  Golden: the reference's own FaceMatcher._aggregate_matches /
_get_best_candidate executed on 400 seeded synthetic tracks (tests/golden/make_golden_flows.py, `tracks/*`):
ties between identities, frames without a match, 0..300 frames per track, four thresholds."""
import os

import numpy as np
import pytest

from facerecognitionpipeline_b200.face_matcher import best_candidate, consensus


def _tracks(golden_dir):
    g = np.load(os.path.join(golden_dir, "flows_cases.npz"))
    return g["tracks/seg"], g["tracks/ids"], g["tracks/scores"], g["tracks/thr"], g["tracks/out"]


def test_host_consensus_equals_reference_methods(golden_dir):
    seg, ids, scores, thr, want = _tracks(golden_dir)
    assert want[:, 0].sum() > 100 and (want[:, 0] == 0).sum() > 100
    for t in range(len(thr)):
        fm = [dict(student_id=int(i), name=f"n{i}", score=float(s))
              for i, s in zip(ids[seg[t]:seg[t + 1]], scores[seg[t]:seg[t + 1]]) if i >= 0]
        w = want[t]
        assert len(fm) == w[5]
        if not fm:
            continue
        a, c = consensus(fm, thr[t]), best_candidate(fm)
        assert (a is not None) == bool(w[0])
        if a:
            assert (a["student_id"], a["confidence"], a["consensus_strength"], a["num_quality_frames"],
                    a["total_frames_evaluated"]) == (w[1], w[2], w[3], w[4], w[5])
        assert (c["student_id"], c["confidence"], c["num_quality_frames"]) == (w[6], w[7], w[8])


@pytest.mark.gpu
def test_device_consensus_is_bit_exact(golden_dir):
    from facerecognitionpipeline_b200.face_matcher import consensus_on_device
    seg, ids, scores, thr, want = _tracks(golden_dir)
    counts = np.diff(seg)
    for th in np.unique(thr):
        sel = np.nonzero(thr == th)[0]
        fr = np.concatenate([np.arange(seg[t], seg[t + 1]) for t in sel]) if len(sel) else np.zeros(0, int)
        # second column = junk: only column 0 (top-1) may be read, through the stride
        ix2 = np.stack([ids[fr], np.full(len(fr), 7)], 1)
        sc2 = np.stack([scores[fr], np.full(len(fr), 0.99, np.float32)], 1)
        got = consensus_on_device(ix2, sc2, counts[sel], th)
        w = want[sel]
        assert np.array_equal(got["recognized"], w[:, 0].astype(np.int32))
        assert np.array_equal(got["winner"], w[:, 1].astype(np.int64))
        assert np.array_equal(got["confidence"], w[:, 2])                   # float64, bit for bit
        assert np.array_equal(got["consensus_strength"], w[:, 3])
        assert np.array_equal(got["num_quality_frames"], w[:, 4].astype(np.int32))
        assert np.array_equal(got["total_frames_evaluated"], w[:, 5].astype(np.int32))
        assert np.array_equal(got["candidate"], w[:, 6].astype(np.int64))
        assert np.array_equal(got["candidate_confidence"], w[:, 7])
        assert np.array_equal(got["candidate_num_quality_frames"], w[:, 8].astype(np.int32))
    assert len(consensus_on_device(np.zeros(0, np.int64), np.zeros(0, np.float32), [], 0.5)) == 0
    with pytest.raises(ValueError):
        consensus_on_device(np.zeros(3, np.int64), np.zeros(3, np.float32), [2], 0.5)


@pytest.mark.gpu
def test_match_tracks_batch_equals_per_track_host_path(tmp_path):
    from oracle import backbone
    from facerecognitionpipeline_b200.face_embedder import FaceEmbedder
    from facerecognitionpipeline_b200.face_matcher import FaceMatcher
    from facerecognitionpipeline_b200.gallery_manager import GalleryManager
    rng = np.random.default_rng(5)
    fe = FaceEmbedder(architecture="ir_50", model_type="adaface", state_dict=backbone.random_state_dict("ir_50", "adaface", seed=0))
    people = [rng.integers(0, 256, (112, 112, 3), dtype=np.uint8) for _ in range(6)]
    gm = GalleryManager(gallery_path=str(tmp_path / "g" / "students.pkl"))
    for i, img in enumerate(people):
        gm.add_student(f"STU{i:04d}", f"P{i}", fe.extract_embeddings_batch([img]))

    def noisy(img, amp):
        return np.clip(img.astype(np.int16) + rng.integers(-amp, amp + 1, img.shape), 0, 255).astype(np.uint8)

    tracks = [[noisy(people[0], 6) for _ in range(5)],
              [noisy(people[1], 6) for _ in range(2)] + [noisy(people[2], 6) for _ in range(2)],     # split vote
              [rng.integers(0, 256, (112, 112, 3), dtype=np.uint8) for _ in range(4)],                 # strangers
              [],
              [noisy(people[3], 6) for _ in range(3)] + [noisy(people[4], 6)]]
    fm = FaceMatcher(similarity_threshold=0.5, embedder=fe, gallery=gm)
    got = fm.match_tracks_batch(tracks, top_k=3)
    assert got[3] is None and len(got) == 5
    for t, crops in enumerate(tracks):
        if not crops:
            continue
        res, _ = fm.match_faces_batch(crops, top_k=3)
        frames = [dict(student_id=m[0][0], name=m[0][1], score=float(m[0][2])) for m in res if m]
        want = consensus(frames, 0.5)
        assert got[t]["recognized"] == (want is not None)
        assert got[t]["frame_matches"] == frames
        if want:
            for key in ("student_id", "name", "confidence", "consensus_strength", "num_quality_frames"):
                assert got[t][key] == want[key]
        else:
            assert got[t]["best_candidate"] == best_candidate(frames)
    assert got[0]["recognized"] and got[0]["student_id"] == "STU0000"
