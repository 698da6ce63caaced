"""Backbone parity: the CUDA path (tcgen05 implicit-GEMM convs, fused epilogues, FC tail) against the fp32
CPU oracle on the same seeded inputs and the same random-init weights.
Bar (BASELINE north star): embedding cosine >= 0.999 (bf16 storage / fp32 accumulate vs fp32)."""
import numpy as np
import pytest
import torch

from oracle import backbone as ob, embedder as oe, emulate, preprocess as op
from facerecognitionpipeline_b200 import _native, weights
from facerecognitionpipeline_b200.face_embedder import FaceEmbedder

pytestmark = pytest.mark.gpu

COS_MIN = 0.999


def _crops(rng, n, S=112):
    import cv2
    return [cv2.GaussianBlur(rng.integers(0, 256, (S, S, 3), dtype=np.uint8), (0, 0), 2.0) for _ in range(n)]


def _cos(a, b):
    return (a * b).sum(1) / (np.linalg.norm(a, axis=1) * np.linalg.norm(b, axis=1))


@pytest.fixture(scope="module")
def ir50():
    sd = ob.random_state_dict("ir_50", "adaface", seed=0)
    return sd, FaceEmbedder("ir_50", model_type="adaface", state_dict=sd), oe.OracleEmbedder("ir_50", "adaface", state_dict=sd)


def test_ir50_32_crops_config1(ir50):
    """BASELINE config 1 embed half: IR-50, 32 synthetic 112x112 crops."""
    sd, fe, orc = ir50
    crops = _crops(np.random.default_rng(0), 32)
    got = fe.extract_embeddings_batch(crops, normalize=True)
    ref = orc.extract_embeddings_batch(crops, normalize=True)
    assert got.shape == (32, 512) and got.dtype == np.float32
    assert np.abs(np.linalg.norm(got, axis=1) - 1).max() < 1e-5
    assert _cos(got, ref).min() >= COS_MIN


def test_api_shapes_and_batch_independence(ir50):
    sd, fe, orc = ir50
    crops = _crops(np.random.default_rng(1), 5)
    assert fe.extract_embeddings_batch([]).size == 0
    one = fe.extract_embedding(crops[0])
    assert one.shape == (512,)
    allb = fe.extract_embeddings_batch(crops, batch_size=2)
    assert np.array_equal(allb[0], one)          # eval mode: a face's embedding does not depend on its batch
    raw = fe.extract_embeddings_batch(crops, normalize=False)
    assert _cos(raw, allb).min() > 0.999999      # adaface output is already unit norm
    assert abs(float(fe.compute_similarity(allb[0], allb[0])) - 1) < 1e-5
    assert fe.aggregate_embeddings(allb, "mean").shape == (512,)


def test_other_input_sizes(ir50):
    """224 crops take the device 2x2 path (== cv2.resize at 2x); other sizes are resized like the reference."""
    sd, fe, orc = ir50
    rng = np.random.default_rng(2)
    crops = _crops(rng, 3, 224) + _crops(rng, 2, 160) + _crops(rng, 1, 112)
    got = fe.extract_embeddings_batch(crops)
    ref = orc.extract_embeddings_batch(crops)
    assert _cos(got, ref).min() >= COS_MIN


def test_device_matches_bf16_emulation(ctx):
    """The device result against a CPU emulation of the SAME folded program with the same bf16 rounding
    points.  bf16 rounding makes the 50-layer stack chaotic in the last bits (an fp32-vs-fp64 run of the
    emulation itself only agrees to cosine ~0.9999), so end to end this can only be as tight as the oracle
    comparison; the tight per-layer check (same inputs, one layer) lives in tests/test_gpu_kernels.py."""
    sd = ob.random_state_dict("ir_50", "adaface", seed=4)
    prog = weights.build_program(sd, "ir_50", "adaface", keep_debug=True)
    prog.load_into(ctx)
    crops = np.stack(_crops(np.random.default_rng(3), 6))
    emb = np.empty((6, 512), np.float32)
    nrm = np.empty((6,), np.float32)
    ctx.frb_embed_host(crops.ctypes.data, 6, 112, _native.FRB_EMBED_L2, emb.ctypes.data, nrm.ctypes.data)
    x = torch.from_numpy(op.preprocess_batch(list(crops), "adaface"))
    emu = emulate.run_program(prog, x, quantize=True).numpy()
    emu_n = emu / np.linalg.norm(emu, axis=1, keepdims=True)
    assert _cos(emb, emu_n).min() >= 0.9995
    np.testing.assert_allclose(nrm, np.linalg.norm(emu, axis=1), rtol=2e-2)


@pytest.mark.parametrize("arch,layout,n", [("ir_101", "adaface", 8), ("ir_50", "iresnet", 6), ("ir_101", "iresnet", 4)])
def test_other_architectures(arch, layout, n):
    sd = ob.random_state_dict(arch, layout, seed=11)
    mt = "adaface" if layout == "adaface" else "arcface"
    fe = FaceEmbedder(arch, model_type=mt, state_dict=sd)
    crops = _crops(np.random.default_rng(12), n)
    for normalize in (True, False):
        got = fe.extract_embeddings_batch(crops, normalize=normalize)
        ref = oe.OracleEmbedder(arch, mt, state_dict=sd).extract_embeddings_batch(crops, normalize=normalize)
        assert _cos(got, ref).min() >= COS_MIN
        if not normalize and layout == "iresnet":   # raw BN1d features: magnitudes must agree too
            assert np.abs(np.linalg.norm(got, axis=1) / np.linalg.norm(ref, axis=1) - 1).max() < 0.03


def test_flip_fusion_config5(ctx):
    """BASELINE config 5 semantics: template = renorm(mean(emb(x), emb(hflip x)))."""
    import cv2
    sd = ob.random_state_dict("ir_50", "iresnet", seed=21)
    prog = weights.build_program(sd, "ir_50", "iresnet")
    prog.load_into(ctx)
    crops = _crops(np.random.default_rng(22), 5)
    stack = np.stack(crops)
    got = np.empty((5, 512), np.float32)
    ctx.frb_embed_host(stack.ctypes.data, 5, 112, _native.FRB_EMBED_FLIP, got.ctypes.data, None)
    orc = oe.OracleEmbedder("ir_50", "arcface", state_dict=sd)
    a = orc.extract_embeddings_batch(crops, normalize=True)
    b = orc.extract_embeddings_batch([cv2.flip(c, 1) for c in crops], normalize=True)
    from oracle import gallery as og
    ref = np.stack([og.aggregate(np.stack([x, y]), "mean") for x, y in zip(a, b)])
    assert _cos(got, ref).min() >= COS_MIN
    assert np.abs(np.linalg.norm(got, axis=1) - 1).max() < 1e-5


def test_large_batch_256_ir101_config2(ctx):
    """BASELINE config 2 shape (IR-101, batch 256): full-size run checked through properties — unit norms,
    batch-composition independence against a batch-8 run of the same faces, and the oracle on 8 faces."""
    sd = ob.random_state_dict("ir_101", "adaface", seed=31)
    fe = FaceEmbedder("ir_101", state_dict=sd, max_batch=256)
    crops = _crops(np.random.default_rng(32), 256)
    big = fe.extract_embeddings_batch(crops)
    assert np.abs(np.linalg.norm(big, axis=1) - 1).max() < 1e-5
    fe.max_batch = 8
    small = fe.extract_embeddings_batch(crops[100:108])
    assert np.array_equal(big[100:108], small)
    ref = oe.OracleEmbedder("ir_101", "adaface", state_dict=sd).extract_embeddings_batch(crops[:8])
    assert _cos(big[:8], ref).min() >= COS_MIN


def test_chunked_prefetch_and_threads(ir50, ctx):
    """Chunked embedding (each chunk's crops are handed to frb_prefetch_host while the previous chunk computes) gives
    the same bytes as one chunk; a prefetch that is never consumed is harmless; and calls from several host threads on
    the shared context (the reference server runs Flask threaded, face_recognition_server.py:1102) serialise on the
    context mutex and return what the sequential calls return."""
    import threading
    sd, fe, orc = ir50
    crops = _crops(np.random.default_rng(41), 100)
    fe.max_batch = 128
    whole = fe.extract_embeddings_batch(crops)
    fe.max_batch = 32                                    # chunks of 32, 32, 32, 4 with prefetch between them
    chunked = fe.extract_embeddings_batch(crops)
    assert np.array_equal(whole, chunked)
    stray = np.ascontiguousarray(np.stack(crops[:7]))
    ctx.frb_prefetch_host(stray.ctypes.data, 7, 112)     # never consumed
    assert np.array_equal(fe.extract_embeddings_batch(crops[:40]), whole[:40])
    out = [None] * 4

    def work(i):
        out[i] = fe.extract_embeddings_batch(crops[i * 25:(i + 1) * 25])

    ts = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert np.array_equal(np.concatenate(out), whole)
    fe.max_batch = 256


def test_device_adaface_embeddings_match_an_independent_onnx_engine(tmp_path):
    """Device IR-50 (AdaFace layout) against OpenCV's dnn module running the same network as an ONNX graph on the host
    (tests/onnx_writer.write_adaface_onnx) - a second, independent judge next to oracle/backbone.py.
    Bar = north_star's: cosine >= 0.999 per face, norm within 1 %."""
    cv2 = pytest.importorskip("cv2")
    from tests.onnx_writer import write_adaface_onnx
    sd = ob.random_state_dict("ir_50", "adaface", seed=3, calibrate=True)
    path = str(tmp_path / "adaface_ir50.onnx")
    write_adaface_onnx(path, sd, "ir_50", batch=6)
    rng = np.random.default_rng(8)
    crops = [cv2.resize(rng.integers(0, 256, (14, 14, 3), dtype=np.uint8), (112, 112), interpolation=cv2.INTER_CUBIC) for _ in range(6)]
    fe = FaceEmbedder("ir_50", model_type="adaface", state_dict=sd)
    dev = fe.extract_embeddings_batch(crops, normalize=False)      # features / ||features|| as the model returns them
    net = cv2.dnn.readNetFromONNX(path)
    net.setInput(np.concatenate([fe.preprocess(c) for c in crops]))
    ref = net.forward()
    ref_unit = ref / np.linalg.norm(ref, axis=1, keepdims=True)
    cos = (ref_unit * dev).sum(1) / np.linalg.norm(dev, axis=1)
    assert cos.min() >= 0.999, cos
