"""Single-kernel parity on the GPU: the tcgen05 GEMM core, the im2col-mode TMA gather, and the implicit-GEMM
convolution with every epilogue variant, each against an independent computation on identical inputs
(torch fp32 on the device / the slow direct CUDA checker / an expected im2col tile built in numpy)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.gpu


def test_tcgen05_gemm_core():
    import gpu_ladder
    res = gpu_ladder.rung_gemm()
    assert all(r["ok"] for r in res), res


def test_im2col_tma_gather():
    import gpu_ladder
    res = gpu_ladder.rung_im2col()
    assert all(r["ok"] for r in res), res


def test_implicit_gemm_conv_all_epilogues():
    """fp32-accumulate conv + {border-bias table, PReLU, identity / strided residual, fused 1x1 shortcut}
    vs torch conv2d (bf16 output rounding => rel 4e-3) and vs the direct CUDA checker."""
    import gpu_ladder
    res = gpu_ladder.rung_conv()
    assert all(r["ok"] for r in res), res
    assert max(r["err_tc_vs_ref"] for r in res) < 4e-3
