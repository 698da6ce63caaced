import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from facerecognitionpipeline_b200 import _native
ctx = _native.Context(0)
g = torch.Generator().manual_seed(0)
slab = torch.randint(-4, 5, (256, 64), generator=g).float()
B = torch.randint(-4, 5, (64, 64), generator=g).float()
sd, Bd = slab.to(torch.bfloat16).cuda(), B.to(torch.bfloat16).cuda()
for mode in (0, 1):
    for j0 in (0, 8, 1, 2, 3, 7, 9, 30, 58, 59, 116, 128):
        out = torch.full((128, 64), float("nan"), device="cuda")
        ctx.frb_debug_shift_mma(sd.data_ptr(), Bd.data_ptr(), j0, mode, out.data_ptr(), None)
        torch.cuda.synchronize()
        ref = slab[j0:j0 + 128] @ B.t()
        bad = (out.cpu() != ref).any(dim=1)
        print(f"mode {mode} j0 {j0:3d}: bad rows {int(bad.sum()):3d}", ("first bad rows " + str([i for i in range(128) if bad[i]][:10])) if bad.any() else "OK", flush=True)
