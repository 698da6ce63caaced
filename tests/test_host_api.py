"""Host-side behaviour of the drop-in modules that needs no GPU: API surface, error conventions,
gallery file layout (pickle + sidecar JSON + backup) and pickle compatibility with the reference's
`gallery_manager.StudentRecord`."""
import json
import os
import pickle
import pickletools
import sys

import numpy as np
import pytest

from facerecognitionpipeline_b200.gallery_manager import GalleryManager, StudentRecord


def _emb(rng, n):
    base = rng.standard_normal(512)
    e = base[None] + 0.3 * rng.standard_normal((n, 512))
    return (e / np.linalg.norm(e, axis=1, keepdims=True)).astype(np.float32)


def test_gallery_crud_and_layout(tmp_path, capsys):
    rng = np.random.default_rng(0)
    path = str(tmp_path / "g" / "students.pkl")
    gm = GalleryManager(gallery_path=path)
    assert gm.get_gallery_embeddings()[1] == [] and gm.get_gallery_embeddings()[0].size == 0
    assert gm.search(np.ones(512, np.float32)) == []           # empty gallery -> [] without touching the GPU
    assert gm.get_statistics() == {"num_students": 0, "total_embeddings": 0, "avg_embeddings_per_student": 0}
    assert gm.add_student("STU0001", "Ann", _emb(rng, 5), metadata={"class": "10A"})
    assert not gm.add_student("STU0001", "Ann", _emb(rng, 5))  # refuses overwrite
    assert gm.add_student("STU0001", "Ann", _emb(rng, 5), overwrite=True)
    assert gm.add_student("STU0002", "Bob", _emb(rng, 1)[0])   # 1-D input reshaped to 1 x 512
    assert gm.get_student("STU0002").embeddings.shape == (1, 512)
    assert np.array_equal(gm.get_student("STU0002").template_embedding, gm.get_student("STU0002").embeddings[0])
    assert not gm.update_embeddings("nope", _emb(rng, 2))
    assert gm.update_embeddings("STU0001", _emb(rng, 3), mode="append") and gm.get_student("STU0001").num_samples == 8
    with pytest.raises(ValueError):
        gm.update_embeddings("STU0001", _emb(rng, 3), mode="bogus")
    assert not gm.delete_student("nope")
    mat, ids = gm.get_gallery_embeddings()
    assert mat.shape == (2, 512) and ids == ["STU0001", "STU0002"]
    gm.save()
    side = json.load(open(path.replace(".pkl", ".json")))
    assert side["num_students"] == 2 and set(side["students"]["STU0001"]) == {
        "student_id", "name", "num_samples", "enrollment_date", "last_updated", "metadata"}
    gm.export_for_backup(str(tmp_path / "bk"), backup_name="adaface_ir_101")
    names = sorted(os.listdir(tmp_path / "bk"))
    assert len(names) == 2 and names[0].startswith("adaface_ir_101_backup_") and names[0].endswith(".json")
    full = json.load(open(tmp_path / "bk" / names[0]))
    assert len(full["students"]["STU0001"]["embeddings"]) == 8 and len(full["students"]["STU0001"]["template_embedding"]) == 512
    gm2 = GalleryManager(gallery_path=path)
    assert list(gm2.students) == ["STU0001", "STU0002"]
    assert gm2.delete_student("STU0002") and len(gm2.students) == 1
    rec = StudentRecord.from_dict(gm.get_student("STU0001").to_dict())
    assert rec.name == "Ann" and rec.embeddings.shape == (8, 512)


def test_pickle_is_reference_compatible(tmp_path):
    """The pickle must name `gallery_manager.StudentRecord` so the reference can load it, and a pickle
    produced under that module name must load here."""
    rng = np.random.default_rng(1)
    path = str(tmp_path / "students.pkl")
    gm = GalleryManager(gallery_path=path)
    gm.add_student("STU0001", "Ann", _emb(rng, 4))
    gm.save()
    ops = [(op.name, arg) for op, arg, _ in pickletools.genops(open(path, "rb").read())]
    strs = [a for n, a in ops if isinstance(a, str)]
    assert "gallery_manager" in strs and "StudentRecord" in strs
    assert not any("facerecognitionpipeline_b200" in s for s in strs)
    assert sys.modules["gallery_manager"].StudentRecord is StudentRecord
    back = pickle.load(open(path, "rb"))
    assert isinstance(back["STU0001"], StudentRecord)


def test_reference_module_loads_our_pickle(tmp_path):
    ref = "/root/reference/gallery_manager.py"
    if not os.path.exists(ref):
        pytest.skip("reference tree not mounted")
    import subprocess
    rng = np.random.default_rng(2)
    path = str(tmp_path / "students.pkl")
    gm = GalleryManager(gallery_path=path)
    gm.add_student("STU0007", "Cy", _emb(rng, 3))
    gm.save()
    code = ("import sys; sys.path.insert(0, '/root/reference'); import gallery_manager as g; "
            f"m = g.GalleryManager(gallery_path={path!r}); s = m.get_student('STU0007'); "
            "assert s.name == 'Cy' and s.embeddings.shape == (3, 512); print('ok')")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=str(tmp_path))
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr


def test_embedder_argument_errors():
    from facerecognitionpipeline_b200.face_embedder import FaceEmbedder
    with pytest.raises(ValueError):
        FaceEmbedder(architecture="ir_18")
    with pytest.raises(ValueError):
        FaceEmbedder(model_type="facenet")
    with pytest.raises(FileNotFoundError):
        FaceEmbedder(architecture="ir_50", model_path="/nonexistent/adaface.ckpt")
    with pytest.raises(FileNotFoundError):
        FaceEmbedder(architecture="ir_50", model_type="arcface")


def test_aligner_template():
    from facerecognitionpipeline_b200.face_recognition import similarity_template
    np.testing.assert_allclose(similarity_template(224)[2], [112.0, 136.64], atol=1e-4)
