"""The enroll_students / embedding_generator flows.

CPU part (no GPU): the host logic of both flows against golden vectors produced by the reference's own
functions (tests/golden/make_golden_flows.py), the file layout they write, the reference's error
dictionaries, and that the device-batched `enroll_from_directory` registers exactly what the
reference's per-student loop would.  A deterministic stand-in embedder / detector / matcher is injected:
the product classes themselves never fall back to CPU.
GPU part (-m gpu): the same flows end to end on the B200 with the real FaceEmbedder / GalleryManager."""
import json
import os
import pickle

import cv2
import numpy as np
import pytest

from facerecognitionpipeline_b200 import embedding_generator as eg
from facerecognitionpipeline_b200 import enroll_students as es
from facerecognitionpipeline_b200.gallery_manager import GalleryManager


class FakeEmbedder:
    """Deterministic unit-norm 512-d 'embedding' of a crop (fixed random projection of the 16x16 thumbnail);
    counts how it was called so the tests can see the batching."""

    def __init__(self):
        self.P = np.random.default_rng(5).standard_normal((16 * 16 * 3, 512)).astype(np.float32)
        self.calls = []

    def extract_embeddings_batch(self, face_images, normalize=True, batch_size=32):
        self.calls.append(len(face_images))
        if len(face_images) == 0:
            return np.array([])
        x = np.stack([cv2.resize(im, (16, 16), interpolation=cv2.INTER_AREA).astype(np.float32).reshape(-1) for im in face_images])
        e = (x - x.mean(1, keepdims=True)) @ self.P
        return (e / np.linalg.norm(e, axis=1, keepdims=True)).astype(np.float32)

    def extract_embedding(self, face_image, normalize=True):
        return self.extract_embeddings_batch([face_image])[0]


class FakeProcessor:
    """Stands in for detect+align: the 'aligned face' is the image itself; files named bad* raise,
    noface* yield nothing, lowq* are flagged invalid."""

    def process_image(self, path, return_all=False):
        name = os.path.basename(path)
        if name.startswith("bad"):
            raise ValueError("cannot decode")
        if name.startswith("noface"):
            return []
        img = cv2.cvtColor(cv2.imread(path), cv2.COLOR_BGR2RGB)
        score = 0.5 + (sum(name.encode()) % 40) / 100.0
        return [{"aligned_face": img, "det_score": score, "quality_metrics": {"blur_score": 200.0},
                 "is_valid": not name.startswith("lowq")}]


class OracleGallery(GalleryManager):
    """GalleryManager whose search runs on the CPU oracle (test infrastructure; the product class matches on the B200)."""

    def search_batch(self, query_embeddings, top_k=5, threshold=0.0):
        from oracle import gallery as og
        mat, ids = self.get_gallery_embeddings()
        q = np.asarray(query_embeddings, np.float32).reshape(-1, 512)
        idx, sc = og.search_batch(np.asarray(mat, np.float32), q, min(top_k, len(ids)))
        res = [[(ids[j], self.students[ids[j]].name, float(s)) for j, s in zip(ir, sr)] for ir, sr in zip(idx, sc)]
        return res, np.array([r[0][2] >= threshold for r in res])


def _person_image(rng, base, S=112):
    img = np.clip(base + rng.normal(0, 12, base.shape), 0, 255).astype(np.uint8)
    return cv2.GaussianBlur(img, (0, 0), 1.0)


def _make_enrollment_tree(root, rng, people=("ann", "bob", "cy"), n=4):
    bases = {}
    for p in people:
        os.makedirs(root / p, exist_ok=True)
        bases[p] = cv2.GaussianBlur(rng.integers(0, 256, (112, 112, 3)).astype(np.float32), (0, 0), 6.0) * 2 - 128
        for i in range(n):
            cv2.imwrite(str(root / p / f"img_{i}.png"), cv2.cvtColor(_person_image(rng, bases[p]), cv2.COLOR_RGB2BGR))
    return bases


# ---------------------------------------------------------------------------------------------- golden
def test_augmentation_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "flows_cases.npz"))
    for i in range(3):
        out = es.augment_face_for_enrollment(g[f"aug/in_{i}"], num_augmentations=8)
        assert len(out) == 8
        assert np.array_equal(np.stack(out), g[f"aug/out_{i}"])        # byte-exact
    img = g["aug/in_0"]
    assert len(es.augment_face_for_enrollment(img, 3)) == 3 and len(es.augment_face_for_enrollment(img, 40)) == 16
    assert eg.augment_face_for_enrollment is es.augment_face_for_enrollment


def test_name_from_filename_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "flows_cases.npz"))
    for fn, want in zip(g["names/in"], g["names/out"]):
        assert eg.EmbeddingGenerator.extract_name_from_filename(str(fn)) == str(want)


# ---------------------------------------------------------------------------------------------- enrollment
def test_enrollment_flow_host_logic(tmp_path):
    rng = np.random.default_rng(0)
    tree = tmp_path / "enroll"
    _make_enrollment_tree(tree, rng)
    os.makedirs(tree / "empty")
    os.makedirs(tree / "dud")
    cv2.imwrite(str(tree / "dud" / "noface_0.png"), np.zeros((112, 112, 3), np.uint8))
    cv2.imwrite(str(tree / "dud" / "lowq_1.png"), np.zeros((112, 112, 3), np.uint8))
    (tree / "dud" / "bad_2.png").write_bytes(b"not a png")
    (tree / "ann" / "notes.txt").write_text("ignored")

    emb = FakeEmbedder()
    gal = OracleGallery(gallery_path=str(tmp_path / "g" / "students.pkl"), aggregation_method="weighted_mean")
    enr = es.StudentEnrollment(min_faces_per_student=2, max_faces_per_student=3, face_processor=FakeProcessor(),
                               embedder=emb, gallery=gal, verbose=False)
    summary = enr.enroll_from_directory(str(tree))
    assert summary["total"] == 5 and summary["successful"] == 3 and summary["failed"] == 2
    by_dir = {os.path.basename(r["directory"]): r for r in summary["results"]}
    assert by_dir["empty"]["info"] == {"error": "no_images"}
    assert by_dir["dud"]["info"] == {"error": "insufficient_faces", "valid_faces": 0, "required": 2}
    assert emb.calls == [3 * 3 * 8]                      # ONE device stream for all students: 3 faces x 8 variants each
    assert [s.student_id for s in gal.students.values()] == ["STU0001", "STU0002", "STU0003"]
    assert [s.name for s in gal.students.values()] == ["ann", "bob", "cy"]
    info = by_dir["bob"]["info"]
    assert set(info) == {"student_id", "name", "num_images", "num_valid_faces", "num_embeddings", "avg_similarity"}
    assert info["num_images"] == 4 and info["num_valid_faces"] == 3 and info["num_embeddings"] == 24
    meta = gal.students["STU0002"].metadata
    assert set(meta) == {"num_images", "num_valid_faces", "num_augmented_faces", "augmentation_per_face", "avg_similarity",
                         "source_directory"}
    assert os.path.exists(tmp_path / "g" / "students.pkl") and os.path.exists(tmp_path / "g" / "students.json")

    # the reference's per-student loop (process_student_directory one at a time) registers the same records
    emb2 = FakeEmbedder()
    gal2 = OracleGallery(gallery_path=str(tmp_path / "g2" / "students.pkl"), aggregation_method="weighted_mean")
    enr2 = es.StudentEnrollment(min_faces_per_student=2, max_faces_per_student=3, face_processor=FakeProcessor(),
                                embedder=emb2, gallery=gal2, verbose=False)
    for d in sorted(os.listdir(tree)):
        enr2.process_student_directory(str(tree / d))
    assert emb2.calls == [24, 24, 24]
    assert list(gal2.students) == list(gal.students)
    for sid in gal.students:
        assert np.array_equal(gal.students[sid].embeddings, gal2.students[sid].embeddings)
        assert np.array_equal(gal.students[sid].template_embedding, gal2.students[sid].template_embedding)
        assert gal.students[sid].metadata["avg_similarity"] == gal2.students[sid].metadata["avg_similarity"]

    # verify_enrollment semantics (enroll_students.py:350-402): first stored embedding retrieves its owner
    rep = enr.verify_enrollment()
    assert rep["rank1_correct"] == rep["total"] == 3 and rep["accuracy"] == 100.0
    # selection knobs
    enr3 = es.StudentEnrollment(min_faces_per_student=1, max_faces_per_student=5, image_indices=[2, 9], face_processor=FakeProcessor(),
                                embedder=FakeEmbedder(), gallery=OracleGallery(gallery_path=str(tmp_path / "g3" / "s.pkl")), verbose=False)
    ok, info = enr3.process_student_directory(str(tree / "ann"), student_id="STU0042")
    assert ok and info["num_images"] == 1 and info["num_embeddings"] == 8 and info["student_id"] == "STU0042"
    with pytest.raises(ValueError):
        enr.enroll_from_directory(str(tmp_path / "missing"))
    assert es.intra_class_similarity(np.eye(4, 512, dtype=np.float32)) == 0.0


# ---------------------------------------------------------------------------------------------- embedding generator
def _make_dataset(tmp_path, rng):
    ds, out = tmp_path / "dataset", tmp_path / "out"
    _make_enrollment_tree(ds / "enrollment" / "one-shot", rng, people=("ann", "bob"), n=1)
    _make_enrollment_tree(ds / "enrollment" / "few-shot", rng, people=("ann", "bob"), n=3)
    pos = out / "probe_labeled" / "positive"
    seg = out / "probe_labeled" / "segmented" / "pose_easy"
    neg = out / "probe_labeled" / "negative"
    for d in (pos, seg, neg):
        os.makedirs(d)
    for nm, S in (("ann_01_a.jpg", 112), ("ann_02_b.png", 150), ("bob_7.jpg", 112)):
        cv2.imwrite(str(pos / nm), rng.integers(0, 256, (S, S, 3), dtype=np.uint8))
    cv2.imwrite(str(seg / "bob_1.png"), rng.integers(0, 256, (112, 112, 3), dtype=np.uint8))
    cv2.imwrite(str(neg / "lfw_X_0001.png"), rng.integers(0, 256, (112, 112, 3), dtype=np.uint8))
    cv2.imwrite(str(neg / "stranger.png"), rng.integers(0, 256, (90, 112, 3), dtype=np.uint8))
    (neg / "broken.png").write_bytes(b"xx")
    return ds, out


def _check_generator_outputs(out, summary, dim_ok):
    d = out / "embeddings" / "adaface_ir_50"
    names = sorted(os.listdir(d))
    assert names == sorted([f"gallery_{t}_{s}.{e}" for t in ("one-shot", "few-shot") for s in ("base", "augmented") for e in ("pkl", "json")]
                           + [f"probe_positive_{s}.{e}" for s in ("segmented", "unsegmented") for e in ("pkl", "json")]
                           + ["probe_negative.pkl", "probe_negative.json", "generation_summary.json"])
    g = pickle.load(open(d / "gallery_few-shot_augmented.pkl", "rb"))
    assert list(g) == ["ann", "bob"] and g["ann"]["embeddings"].shape == (24, 512) and g["ann"]["num_images"] == 3
    assert set(g["ann"]) == {"embeddings", "num_images", "num_embeddings", "image_files", "enrollment_type", "augmented"}
    assert g["ann"]["augmented"] is True and g["ann"]["enrollment_type"] == "few-shot"
    gj = json.load(open(d / "gallery_one-shot_base.json"))
    assert len(gj["bob"]["embeddings"]) == 1 and len(gj["bob"]["embeddings"][0]) == 512
    pp = pickle.load(open(d / "probe_positive_unsegmented.pkl", "rb"))
    assert list(pp) == ["all"] and sorted(pp["all"]) == ["ann", "bob"]
    assert pp["all"]["ann"]["embeddings"].shape == (2, 512) and pp["all"]["ann"]["filenames"] == ["ann_01_a.jpg", "ann_02_b.png"]
    ps = pickle.load(open(d / "probe_positive_segmented.pkl", "rb"))
    assert list(ps) == ["pose_easy"]
    pn = pickle.load(open(d / "probe_negative.pkl", "rb"))
    assert pn["lfw"]["filenames"] == ["lfw_X_0001.png"] and pn["real"]["filenames"] == ["stranger.png"]
    assert pn["real"]["embeddings"].shape == (1, 512)
    assert summary["gallery"] == {"one_shot_base_persons": 2, "one_shot_augmented_persons": 2, "few_shot_base_persons": 2,
                                  "few_shot_augmented_persons": 2}
    assert summary["probe_negative"] == {"real_images": 1, "lfw_images": 1}
    assert summary["probe_positive"]["segmented_categories"] == ["pose_easy"]
    assert json.load(open(d / "generation_summary.json"))["model_name"] == "adaface_ir_50"
    dim_ok(g["ann"]["embeddings"])


def test_embedding_generator_layout(tmp_path):
    rng = np.random.default_rng(1)
    ds, out = _make_dataset(tmp_path, rng)
    emb = FakeEmbedder()
    gen = eg.EmbeddingGenerator("adaface", "ir_50", dataset_root=ds, output_root=out, embedder=emb,
                                face_processor=FakeProcessor(), verbose=False)
    summary = gen.generate_all_embeddings()
    _check_generator_outputs(out, summary, lambda e: None)
    assert emb.calls == [2, 16, 6, 48, 3, 1, 2]      # one device stream per tree, not one forward pass per image
    # probes are resized like the reference (cv2.resize default) before embedding
    crop = gen._load_crop(out / "probe_labeled" / "positive" / "ann_02_b.png")
    ref = cv2.resize(cv2.cvtColor(cv2.imread(str(out / "probe_labeled" / "positive" / "ann_02_b.png")), cv2.COLOR_BGR2RGB), (112, 112))
    assert np.array_equal(crop, ref)
    no_det = eg.EmbeddingGenerator("adaface", "ir_50", dataset_root=ds, output_root=tmp_path / "o2", embedder=emb, verbose=False)
    with pytest.raises(ImportError):
        no_det.process_gallery_enrollment("one-shot")
    assert no_det.process_probe_negative() == {}        # missing tree -> {} like the reference
    assert eg.to_serializable({"a": np.float32(1.5), "b": [np.int64(2), np.zeros(2)]}) == {"a": 1.5, "b": [2, [0.0, 0.0]]}


# ---------------------------------------------------------------------------------------------- on the B200
@pytest.mark.gpu
def test_flows_end_to_end_on_device(tmp_path):
    from oracle import backbone as ob
    from facerecognitionpipeline_b200.face_embedder import FaceEmbedder
    rng = np.random.default_rng(2)
    fe = FaceEmbedder("ir_50", model_type="adaface", state_dict=ob.random_state_dict("ir_50", "adaface", seed=0))
    tree = tmp_path / "enroll"
    _make_enrollment_tree(tree, rng, people=("ann", "bob", "cy", "di"), n=3)
    gal = GalleryManager(gallery_path=str(tmp_path / "g" / "students.pkl"), aggregation_method="weighted_mean")
    enr = es.StudentEnrollment(min_faces_per_student=1, max_faces_per_student=5, face_processor=FakeProcessor(), embedder=fe,
                               gallery=gal, verbose=False)
    summary = enr.enroll_from_directory(str(tree))
    assert summary["successful"] == 4 and summary["gallery_stats"]["total_embeddings"] == 4 * 3 * 8
    rep = enr.verify_enrollment()                      # batched device search
    assert rep["rank1_correct"] == 4
    # device-batched enrollment == the reference's per-student loop, bit for bit
    gal2 = GalleryManager(gallery_path=str(tmp_path / "g2" / "students.pkl"), aggregation_method="weighted_mean")
    enr2 = es.StudentEnrollment(min_faces_per_student=1, max_faces_per_student=5, face_processor=FakeProcessor(), embedder=fe,
                                gallery=gal2, verbose=False)
    for d in sorted(os.listdir(tree)):
        enr2.process_student_directory(str(tree / d))
    for sid in gal.students:
        assert np.array_equal(gal.students[sid].embeddings, gal2.students[sid].embeddings)
    ds, out = _make_dataset(tmp_path, rng)
    gen = eg.EmbeddingGenerator("adaface", "ir_50", dataset_root=ds, output_root=out, embedder=fe, face_processor=FakeProcessor(),
                                verbose=False)
    summary = gen.generate_all_embeddings()
    _check_generator_outputs(out, summary, lambda e: np.testing.assert_allclose(np.linalg.norm(e, axis=1), 1.0, atol=1e-5))
