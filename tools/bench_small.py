#!/usr/bin/env python
"""Embed latency at small batches (the reference's single-image / server path): IR-101, device-resident input."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from facerecognitionpipeline_b200 import _native, weights
ctx = _native.Context(0)
dev = torch.device("cuda", 0)
prog = weights.build_program(weights.random_init_state_dict("ir_101", "adaface", seed=0), "ir_101", "adaface")
prog.load_into(ctx)
flags = _native.FRB_EMBED_L2 | _native.FRB_EMBED_RENORM
st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
for B in [int(a) for a in sys.argv[1:]] or [1, 8, 32, 64, 128]:
    x = torch.randn((B, 112, 112, 3), device=dev).to(torch.bfloat16)
    emb = torch.empty((B, 512), dtype=torch.float32, device=dev)
    f = lambda: ctx.frb_embed(x.data_ptr(), B, flags, emb.data_ptr(), None, None, st)
    for _ in range(5): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(30): f()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 30
    print(f"B={B:4d}: {ms:7.3f} ms per embed  {B / ms * 1e3:9.0f} faces/s  checksum {float(emb.abs().sum()):.4f}", flush=True)
