"""Oracle backbone structure + the product's BN folding, all on CPU.
The backbone oracle is 'parity unpinned' (see oracle/backbone.py header): these tests pin what CAN be pinned —
published FLOP / parameter counts, state-dict key names, and that the folded device program is the
same function as the unfolded oracle."""
import numpy as np
import pytest
import torch

from oracle import backbone as ob, emulate
from facerecognitionpipeline_b200 import weights


def test_flops_match_survey():
    assert abs(ob.flops_per_face("ir_101", "adaface") / 1e9 - 24.154) < 1e-3
    assert abs(ob.flops_per_face("ir_50", "adaface") / 1e9 - 12.593) < 1e-3
    assert abs(ob.flops_per_face("ir_101", "iresnet") / 1e9 - 24.179) < 1e-3
    assert abs(ob.flops_per_face("ir_50", "iresnet") / 1e9 - 12.619) < 1e-3


@pytest.mark.parametrize("arch,units,params_m", [("ir_50", 24, 43.59), ("ir_101", 49, 65.15)])
def test_adaface_state_dict_layout(arch, units, params_m):
    sd = ob.random_state_dict(arch, "adaface", 0, calibrate=False)
    assert len(ob.unit_specs(arch)) == units
    n = sum(v.numel() for k, v in sd.items() if v.ndim > 0 and "running" not in k)
    assert abs(n / 1e6 - params_m) < 0.01
    for k in ("input_layer.0.weight", "input_layer.2.weight", "body.0.res_layer.1.weight", "body.0.res_layer.3.weight",
              f"body.{units - 1}.res_layer.5.running_var", "output_layer.3.weight", "output_layer.4.running_mean"):
        assert k in sd, k
    assert "output_layer.4.weight" not in sd            # BN1d(affine=False)
    assert "body.0.shortcut_layer.0.weight" not in sd   # 64->64 stride 2: MaxPool2d(1,2) shortcut
    assert "body.3.shortcut_layer.0.weight" in sd       # first unit of stage 2: Conv1x1+BN


def test_iresnet_has_downsample_in_every_stage():
    sd = ob.random_state_dict("ir_50", "iresnet", 0, calibrate=False)
    for st in range(1, 5):
        assert f"layer{st}.0.downsample.0.weight" in sd
    assert "features.weight" in sd and "fc.bias" in sd


def test_unknown_arch_raises():
    with pytest.raises(ValueError):
        ob.unit_specs("ir_18")
    with pytest.raises(ValueError):
        weights.build_program({}, "ir_18", "adaface")
    with pytest.raises(ValueError):
        weights.build_program({}, "ir_50", "onnx")


@pytest.mark.parametrize("layout", ["adaface", "iresnet"])
def test_folded_program_equals_oracle(layout):
    """fp32 execution of the folded program (bf16-rounded weights) vs the unfolded oracle: cosine >= 0.9999;
    with bf16 activation rounding at every layer boundary (what the device stores): >= 0.999."""
    arch = "ir_50"
    sd = ob.random_state_dict(arch, layout, 3)
    g = torch.Generator().manual_seed(9)
    x = (torch.randint(0, 256, (3, 3, 112, 112), generator=g).float() / 255 - 0.5) / 0.5
    ref = ob.forward(sd, x, arch, layout)
    ref = ref[0] * ref[1] if layout == "adaface" else ref
    prog = weights.build_program(sd, arch, layout, keep_debug=True)
    assert len(prog.layers) == 1 + 2 * 24 + 1
    cos32 = torch.nn.functional.cosine_similarity(emulate.run_program(prog, x, quantize=False), ref).min().item()
    cos16 = torch.nn.functional.cosine_similarity(emulate.run_program(prog, x, quantize=True), ref).min().item()
    assert cos32 >= 0.9999, cos32
    assert cos16 >= 0.999, cos16


def test_border_bias_table_cases():
    T = torch.arange(9, dtype=torch.float64)[:, None] + 1  # tap t contributes t+1
    tab = weights.border_bias_table(T, torch.zeros(1, dtype=torch.float64))
    assert tab[4, 0] == 45                      # interior: all nine taps
    assert tab[0, 0] == 5 + 6 + 8 + 9           # top-left corner keeps taps (1,1),(1,2),(2,1),(2,2)
    assert tab[8, 0] == 1 + 2 + 4 + 5           # bottom-right corner
    assert tab[1, 0] == 4 + 5 + 6 + 7 + 8 + 9   # top edge drops r = 0


@pytest.mark.parametrize("arch", ["ir_50", "ir_101"])
def test_adaface_forward_equals_an_independent_onnx_engine(tmp_path, arch):
    """The oracle's AdaFace forward against OpenCV's dnn module running the same network written out as an ONNX graph
    (tests/onnx_writer.write_adaface_onnx): two independent implementations of conv padding / stride placement, the
    MaxPool(1, stride) shortcut, BatchNorm epsilon, PReLU broadcasting, NCHW flatten order and the affine-free last
    BatchNorm.  What stays unpinned is whether this graph IS mk-minchul/AdaFace's net.py (un-vendored, no weights in
    the reference); the arithmetic of the graph as restated is pinned to 1e-4 relative."""
    cv2 = pytest.importorskip("cv2")
    if not hasattr(cv2, "dnn"):
        pytest.skip("OpenCV without dnn")
    import numpy as np
    from tests.onnx_writer import write_adaface_onnx
    sd = ob.random_state_dict(arch, "adaface", seed=3, calibrate=True)
    path = str(tmp_path / "adaface.onnx")
    write_adaface_onnx(path, sd, arch, batch=2)
    x = np.random.default_rng(1).standard_normal((2, 3, 112, 112)).astype(np.float32)
    net = cv2.dnn.readNetFromONNX(path)
    net.setInput(x)
    ref = net.forward()
    feat, norm = ob.forward(sd, torch.from_numpy(x), arch, "adaface")
    got = (feat * norm).numpy()
    assert np.abs(ref - got).max() <= 1e-4 * np.abs(ref).max()
    assert np.allclose(np.linalg.norm(ref, axis=1), norm.flatten().numpy(), rtol=1e-5)
