import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from facerecognitionpipeline_b200 import _native
ctx = _native.Context(0)
out = (C.c_longlong * 2)()
for N in (64, 128, 256):
    for iters in (1024,):
        for mode in (0, 2 + 9):
            ctx.frb_debug_mma_rate(N, iters, mode, out); ctx.frb_debug_mma_rate(N, iters, mode, out)
            print(f"1-CTA 128x{N:3d}x16 mode {mode:2d}: issue {out[0]/iters:7.1f} cyc/MMA  complete {out[1]/iters:7.1f} cyc/MMA", flush=True)
        for off in (0, 9, 59):
            ctx.frb_debug_mma2_rate(N, iters, off, out); ctx.frb_debug_mma2_rate(N, iters, off, out)
            print(f"pair  256x{N:3d}x16 A row offset {off:2d}: issue {out[0]/iters:7.1f} cyc/MMA  complete {out[1]/iters:7.1f} cyc/MMA", flush=True)
