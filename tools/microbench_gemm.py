#!/usr/bin/env python
"""Micro-benchmarks that separate the bounds of the tcgen05 kernels (run on the GPU box)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from facerecognitionpipeline_b200 import _native
from facerecognitionpipeline_b200._native import LayerDesc

ctx = _native.Context(0)
dev = torch.device("cuda", 0)

def timeit(fn, reps=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3  # us

def gemm(M, N, K):
    A = torch.randn(M, K, device=dev).to(torch.bfloat16)
    B = torch.randn(N, K, device=dev).to(torch.bfloat16)
    Cc = torch.empty(M, N, device=dev, dtype=torch.float32)
    us = timeit(lambda: ctx.frb_debug_gemm(A.data_ptr(), B.data_ptr(), M, N, K, 1, Cc.data_ptr(), None))
    print(f"gemm tiled 1cta  M={M:6d} N={N:3d} K={K:5d}: {us:8.1f} us  {2*M*N*K/us/1e6:7.1f} TF", flush=True)
    us2 = timeit(lambda: torch.matmul(A, B.t()))
    print(f"   torch.matmul (cuBLAS) same shape: {us2:8.1f} us  {2*M*N*K/us2/1e6:7.1f} TF", flush=True)

def conv(Bn, H, Cin, Cout, stride=1, mode="pair"):
    L = LayerDesc()
    L.op = 1; L.cin = Cin; L.cout = Cout; L.hin = H; L.win = H; L.ksize = 3; L.stride = stride; L.pad = 1
    L.in_buf = 0; L.out_buf = 1; L.sc_buf = -1; L.res_buf = -1; L.bias_cases = 9; L.has_prelu = 1
    P = (H + 2 - 3) // stride + 1
    x = torch.randn(Bn, H, H, Cin, device=dev).to(torch.bfloat16)
    w = (torch.randn(Cout, 9 * Cin, device=dev) / (9 * Cin) ** 0.5).to(torch.bfloat16)
    bias = torch.randn(9, Cout, device=dev); slope = torch.rand(Cout, device=dev)
    o = torch.empty(Bn, P, P, Cout, device=dev, dtype=torch.bfloat16)
    us = timeit(lambda: ctx.frb_debug_conv(L, Bn, x.data_ptr(), None, None, w.data_ptr(), bias.data_ptr(), slope.data_ptr(), o.data_ptr(), 0, None))
    fl = 2.0 * Bn * P * P * Cout * 9 * Cin
    print(f"conv {mode} B={Bn:4d} {H}x{H} {Cin}->{Cout} s{stride}: {us:8.1f} us  {fl/us/1e6:7.1f} TF  tiles/SM={Bn*P*P/128/148:.2f}", flush=True)

which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "gemm"):
    gemm(50176, 256, 2304); gemm(56832, 256, 2304); gemm(56832, 256, 9216); gemm(18944, 256, 2304); gemm(8192, 8192, 8192)
    gemm(200704, 128, 1152); gemm(802816 // 4, 64, 576)
if which in ("all", "conv"):
    mode = "pair" if os.environ.get("FRB_CONV_MODE", "2") == "2" else "mcast"
    conv(256, 14, 256, 256, mode=mode); conv(290, 14, 256, 256, mode=mode); conv(97, 14, 256, 256, mode=mode); conv(1024, 14, 256, 256, mode=mode)
    conv(256, 28, 128, 128, mode=mode); conv(256, 56, 64, 64, mode=mode); conv(256, 7, 512, 512, mode=mode); conv(1024, 7, 512, 512, mode=mode)
if which == "one":  # one B H Cin Cout [stride]
    a = [int(x) for x in sys.argv[2:]]
    conv(a[0], a[1], a[2], a[3], a[4] if len(a) > 4 else 1, mode="one")
if which == "slab56":
    conv(256, 56, 64, 64, mode="slab")
if which == "slab28":
    conv(256, 28, 128, 128, mode="slab")
