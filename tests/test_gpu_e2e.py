"""End-to-end parity through the reference-facing API (BASELINE config 1): IR-50 embedding of 32 synthetic
aligned crops + cosine match against a 100-identity gallery via FaceMatcher.match_single_face, against the
oracle: top-1 identity and accept/reject 100 % identical, top-k indices identical."""
import numpy as np
import pytest

from oracle import backbone as ob, embedder as oe, gallery as og
from facerecognitionpipeline_b200.face_embedder import FaceEmbedder
from facerecognitionpipeline_b200.face_matcher import FaceMatcher
from facerecognitionpipeline_b200.gallery_manager import GalleryManager

pytestmark = pytest.mark.gpu


def _crops(rng, n):
    import cv2
    return [cv2.GaussianBlur(rng.integers(0, 256, (112, 112, 3), dtype=np.uint8), (0, 0), 2.0) for _ in range(n)]


def test_config1_match_single_face_loop(tmp_path):
    rng = np.random.default_rng(0)
    sd = ob.random_state_dict("ir_50", "adaface", seed=0)
    orc = oe.OracleEmbedder("ir_50", "adaface", state_dict=sd)
    enrolled = _crops(rng, 40)
    probes = enrolled[:24] + _crops(rng, 8)                      # 24 genuine + 8 impostors = 32 crops
    # gallery of 100 identities built from ORACLE embeddings: 40 enrolled faces + 60 random unit vectors
    E = orc.extract_embeddings_batch(enrolled)
    R = rng.standard_normal((60, 512)).astype(np.float32)
    R /= np.linalg.norm(R, axis=1, keepdims=True)
    G = np.vstack([E, R]).astype(np.float32)
    gm = GalleryManager(gallery_path=str(tmp_path / "students.pkl"))
    for i, row in enumerate(G):
        gm.add_student(f"STU{i:04d}", f"Student {i}", row)
    gm.save()
    thr = 0.5
    fm = FaceMatcher(gallery_path=str(tmp_path / "students.pkl"), similarity_threshold=thr, architecture="ir_50",
                     embedder=FaceEmbedder("ir_50", state_dict=sd))
    ref_emb = orc.extract_embeddings_batch(probes)
    eidx, esc = og.search_batch(G, ref_emb, 5)
    # the oracle's own decision margins must exceed the embedding tolerance, else the case is ill-posed
    assert (esc[:, 0] - esc[:, 1]).min() > 0.02
    assert np.abs(esc[:, 0] - thr).min() > 0.02
    for p, crop in enumerate(probes):
        res = fm.match_single_face(crop, top_k=5)
        assert [r[0] for r in res][0] == f"STU{eidx[p, 0]:04d}"                     # top-1 identity
        assert (res[0][2] >= thr) == (esc[p, 0] >= thr)                             # accept / reject
        assert abs(res[0][2] - esc[p, 0]) < 5e-3
        assert isinstance(res[0][2], float) and res[0][1] == f"Student {eidx[p, 0]}"
    batch, accept = fm.match_faces_batch(probes, top_k=5)
    assert [b[0][0] for b in batch] == [f"STU{i:04d}" for i in eidx[:, 0]]
    assert np.array_equal(accept, esc[:, 0] >= thr)
    assert accept[:24].all() and not accept[24:].any()
    # top-k: identical index lists when matching the SAME embeddings (match parity proper is test_gpu_match)
    res, _ = gm.search_batch(ref_emb, 5)
    assert [[r[0] for r in row] for row in res] == [[f"STU{i:04d}" for i in row] for row in eidx]
