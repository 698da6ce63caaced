import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from facerecognitionpipeline_b200 import _native
ctx = _native.Context(0)
out = (C.c_longlong * 2)()
for mode in (0, 2 + 1, 2 + 2, 2 + 4, 2 + 8, 2 + 9, 2 + 58, 2 + 30):
    for N in (64, 128, 256):
        for iters in (64, 1024):
            ctx.frb_debug_mma_rate(N, iters, mode, out)
            ctx.frb_debug_mma_rate(N, iters, mode, out)
            print(f"mode {mode} N={N:3d} iters={iters:5d}: issue {out[0]/iters:7.1f} cyc/MMA  complete {out[1]/iters:7.1f} cyc/MMA", flush=True)
