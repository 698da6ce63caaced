"""The C-ABI library builds, loads on a CPU-only box and exports every symbol include/frb200.h declares.
No compute calls here (no GPU)."""
import os
import re

import pytest

from facerecognitionpipeline_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "frb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(frb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    if not _native.LIB_PATH.exists():
        _native.build()
    lib = _native.lib()
    syms = header_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/frb200.h but not exported"
    assert set(syms) == set(_native.SIGNATURES), set(syms) ^ set(_native.SIGNATURES)


def test_struct_sizes_match_header():
    import ctypes as C
    assert C.sizeof(_native.WarpJob) == 8 + 16 + 48
    assert C.sizeof(_native.LayerDesc) == 23 * 4 + 4 + 4 * 8  # 23 int32 + pad + 4 int64
    assert C.sizeof(_native.TrackResult) == 56                # frb_track_result: i64, 2 x f64, 2 x i32, i64, f64, 2 x i32


def test_no_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_native.NativeError):
        _native.Context(0)
    from facerecognitionpipeline_b200.face_embedder import FaceEmbedder
    from oracle import backbone
    with pytest.raises(_native.NativeError):
        FaceEmbedder("ir_50", state_dict=backbone.random_state_dict("ir_50", "adaface", 0, calibrate=False))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "facerecognitionpipeline_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn
