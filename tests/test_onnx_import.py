"""ArcFace ONNX import (reference face_embedder.py:64-88 loads `arcface_ir{50,101}_ms1mv3.onnx`): a synthetic export
written by tests/onnx_writer.py must come back as the iresnet state dict it was written from, and build the same
device program.  CPU tests: reader + mapping + program equality; GPU test: same embeddings as the state-dict path."""
import numpy as np
import pytest
import torch

from facerecognitionpipeline_b200 import onnx_import, weights
from oracle import backbone as ob
from tests.onnx_writer import write_iresnet_onnx


@pytest.fixture(scope="module")
def sd50():
    return ob.random_state_dict("ir_50", "iresnet", seed=5, calibrate=False)


def test_unfolded_export_round_trips_exactly(tmp_path, sd50):
    path = str(tmp_path / "arcface_ir50.onnx")
    write_iresnet_onnx(path, sd50, weights.UNITS["ir_50"], fold_conv_bn=False)
    sd, arch = onnx_import.onnx_to_iresnet_state_dict(path)
    assert arch == "ir_50"
    for k, v in sd50.items():
        if k.endswith("num_batches_tracked"):
            continue
        assert k in sd, k
        assert torch.equal(sd[k].float().reshape(v.shape), v.float()), k
    a = weights.build_program(sd50, "ir_50", "iresnet")
    b = weights.build_program(sd, "ir_50", "iresnet")
    assert a.blob == b.blob and len(a.layers) == len(b.layers)            # the device program is byte-identical


@pytest.mark.parametrize("gemm", [True, False])
def test_folded_export_builds_the_same_program(tmp_path, sd50, gemm):
    """Exporter-side Conv+BN folding (numeric initializer names, Conv bias) and MatMul+Add instead of Gemm."""
    path = str(tmp_path / "folded.onnx")
    write_iresnet_onnx(path, sd50, weights.UNITS["ir_50"], fold_conv_bn=True, gemm=gemm)
    sd, arch = onnx_import.onnx_to_iresnet_state_dict(path)
    assert arch == "ir_50"
    a = weights.build_program(sd50, "ir_50", "iresnet", keep_debug=True)
    b = weights.build_program(sd, "ir_50", "iresnet", keep_debug=True)
    for da, db in zip(a.debug, b.debug):
        for k in da:
            if isinstance(da[k], torch.Tensor):
                # the exporter rounded the folded weights to f32 once more: agreement to bf16 / f32 rounding
                tol = 2 ** -7 if k in ("w", "w1", "w2") else 1e-5
                assert torch.allclose(da[k], db[k], rtol=tol, atol=1e-5 * float(da[k].abs().max())), k


def test_ir101_unit_count_and_errors(tmp_path):
    sd = ob.random_state_dict("ir_101", "iresnet", seed=6, calibrate=False)
    path = str(tmp_path / "r100.onnx")
    write_iresnet_onnx(path, sd, weights.UNITS["ir_101"])
    got, arch = onnx_import.onnx_to_iresnet_state_dict(path)
    assert arch == "ir_101" and torch.equal(got["layer3.29.conv2.weight"], sd["layer3.29.conv2.weight"])
    bad = str(tmp_path / "bad.onnx")
    write_iresnet_onnx(bad, sd, [3, 13, 30, 2])                          # 48 units: neither iresnet50 nor iresnet100
    with pytest.raises(ValueError, match="residual units"):
        onnx_import.onnx_to_iresnet_state_dict(bad)
    open(str(tmp_path / "junk.onnx"), "wb").write(b"\x0a\x03abc")
    with pytest.raises(ValueError):
        onnx_import.onnx_to_iresnet_state_dict(str(tmp_path / "junk.onnx"))


@pytest.mark.gpu
def test_arcface_embedder_from_onnx_equals_state_dict_path(tmp_path, sd50):
    from facerecognitionpipeline_b200.face_embedder import FaceEmbedder
    path = str(tmp_path / "arcface_ir50_ms1mv3.onnx")
    write_iresnet_onnx(path, sd50, weights.UNITS["ir_50"])
    rng = np.random.default_rng(3)
    crops = [rng.integers(0, 256, (112, 112, 3), dtype=np.uint8) for _ in range(5)]
    a = FaceEmbedder("ir_50", model_path=path, model_type="arcface").extract_embeddings_batch(crops, normalize=False)
    b = FaceEmbedder("ir_50", model_type="arcface", state_dict=sd50).extract_embeddings_batch(crops, normalize=False)
    assert np.array_equal(a, b)                                           # same program bytes -> same bits
    with pytest.raises(ValueError, match="ir_50"):
        FaceEmbedder("ir_101", model_path=path, model_type="arcface")     # the file holds the other depth


def _opencv_dnn():
    try:
        import cv2
        cv2.dnn.readNetFromONNX
        return cv2
    except Exception:                                   # pragma: no cover - the image ships opencv with dnn
        pytest.skip("OpenCV dnn module not available")


@pytest.fixture(scope="module")
def sd50_cal():
    return ob.random_state_dict("ir_50", "iresnet", seed=5, calibrate=True)      # BN statistics calibrated: activations stay O(1)


@pytest.mark.parametrize("fold", [False, True])
def test_oracle_iresnet_equals_an_independent_onnx_engine(tmp_path, sd50_cal, fold):
    """Pin for the ArcFace path.  The reference runs `arcface_ir*_ms1mv3.onnx` through an ONNX runtime
    (face_embedder.py:64-88, insightface model_zoo), i.e. its arithmetic is "whatever the ONNX operator semantics say
    for this graph".  OpenCV's dnn module is an independent implementation of those semantics (Conv padding / strides,
    BatchNormalization epsilon, PRelu slope broadcasting, Gemm transB, Flatten): it must agree with oracle/backbone.py's
    iresnet forward on the same export - kept-BN and exporter-folded flavours."""
    cv2 = _opencv_dnn()
    path = str(tmp_path / "arcface_ir50.onnx")
    write_iresnet_onnx(path, sd50_cal, weights.UNITS["ir_50"], fold_conv_bn=fold, batch=2)
    x = np.random.default_rng(0).standard_normal((2, 3, 112, 112)).astype(np.float32)
    net = cv2.dnn.readNetFromONNX(path)
    net.setInput(x)
    ref = net.forward()
    with torch.no_grad():
        got = ob.forward(sd50_cal, torch.from_numpy(x), "ir_50", "iresnet").numpy()
    assert ref.shape == got.shape == (2, 512)
    assert np.abs(ref - got).max() <= 1e-4 * np.abs(ref).max()           # fp32 engines, different summation orders
    cos = (ref * got).sum(1) / np.linalg.norm(ref, axis=1) / np.linalg.norm(got, axis=1)
    assert cos.min() > 0.999999


@pytest.mark.gpu
def test_device_arcface_embeddings_match_the_independent_onnx_engine(tmp_path, sd50_cal):
    """Same pin, one step further: crops -> FaceEmbedder(model_path=<onnx>) on the device against OpenCV dnn running
    the file on the host, fed with FaceEmbedder.preprocess (the reference's own preprocessing, face_embedder.py:93-110).
    Tolerance = north_star's embedding bar (cosine >= 0.999; the device computes in bf16)."""
    cv2 = _opencv_dnn()
    from facerecognitionpipeline_b200.face_embedder import FaceEmbedder
    path = str(tmp_path / "arcface_ir50_ms1mv3.onnx")
    write_iresnet_onnx(path, sd50_cal, weights.UNITS["ir_50"], batch=4)
    rng = np.random.default_rng(4)
    base = rng.integers(0, 256, (4, 14, 14, 3), dtype=np.uint8)
    crops = [cv2.resize(b, (112, 112), interpolation=cv2.INTER_CUBIC) for b in base]     # smooth, face-like spectra
    emb = FaceEmbedder("ir_50", model_path=path, model_type="arcface")
    dev = emb.extract_embeddings_batch(crops, normalize=False)
    net = cv2.dnn.readNetFromONNX(path)
    net.setInput(np.concatenate([emb.preprocess(c) for c in crops]))
    ref = net.forward()
    cos = (ref * dev).sum(1) / np.linalg.norm(ref, axis=1) / np.linalg.norm(dev, axis=1)
    assert cos.min() >= 0.999, cos
