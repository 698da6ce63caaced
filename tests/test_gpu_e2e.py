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
    thr = 0.85   # genuine probes score ~1.0, impostors (other noise crops) ~0.6-0.72 with random-init weights
    fm = FaceMatcher(gallery_path=str(tmp_path / "students.pkl"), similarity_threshold=thr, architecture="ir_50",
                     embedder=FaceEmbedder("ir_50", state_dict=sd))
    ref_emb = orc.extract_embeddings_batch(probes)
    eidx, esc = og.search_batch(G, ref_emb, 5)
    # accept/reject must be well-posed for every probe; top-1 identity for every probe whose oracle margin
    # (top-1 minus top-2) exceeds the embedding tolerance — random-init nets map unrelated noise crops to
    # near-equidistant embeddings, so impostors' "identity" is decided by margins of ~1e-3 and is ill-posed.
    margin = esc[:, 0] - esc[:, 1]
    well_posed = margin > 0.02
    assert well_posed[:24].all()
    assert np.abs(esc[:, 0] - thr).min() > 0.05
    for p, crop in enumerate(probes):
        res = fm.match_single_face(crop, top_k=5)
        if well_posed[p]:
            assert res[0][0] == f"STU{eidx[p, 0]:04d}"                              # top-1 identity
            assert res[0][1] == f"Student {eidx[p, 0]}"
        assert (res[0][2] >= thr) == (esc[p, 0] >= thr)                             # accept / reject
        assert abs(res[0][2] - esc[p, 0]) < 5e-3 and isinstance(res[0][2], float)
    batch, accept = fm.match_faces_batch(probes, top_k=5)
    assert [b[0][0] for b, w in zip(batch, well_posed) if w] == [f"STU{i:04d}" for i, w in zip(eidx[:, 0], well_posed) if w]
    assert np.array_equal(accept, esc[:, 0] >= thr)
    assert accept[:24].all() and not accept[24:].any()
    # top-k: identical index lists when matching the SAME embeddings (match parity proper is test_gpu_match)
    res, _ = gm.search_batch(ref_emb, 5)
    assert [[r[0] for r in row] for row in res] == [[f"STU{i:04d}" for i in row] for row in eidx]
