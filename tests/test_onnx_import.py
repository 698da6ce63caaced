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
