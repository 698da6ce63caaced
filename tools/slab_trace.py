import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
tr = torch.zeros(148 * 128, dtype=torch.int64, device="cuda")
os.environ["FRB_SLAB_TRACE"] = hex(tr.data_ptr())
os.environ["FRB_SLAB"] = "1"
from facerecognitionpipeline_b200 import _native
from facerecognitionpipeline_b200._native import LayerDesc
ctx = _native.Context(0)
dev = torch.device("cuda", 0)
def conv(Bn, H, Cin, Cout):
    L = LayerDesc()
    L.op = 1; L.cin = Cin; L.cout = Cout; L.hin = H; L.win = H; L.ksize = 3; L.stride = 1; L.pad = 1
    L.in_buf = 0; L.out_buf = 1; L.sc_buf = -1; L.res_buf = -1; L.bias_cases = 9; L.has_prelu = 1
    x = torch.randn(Bn, H, H, Cin, device=dev).to(torch.bfloat16)
    w = (torch.randn(Cout, 9 * Cin, device=dev) / (9 * Cin) ** 0.5).to(torch.bfloat16).repeat(74, 1).contiguous()
    bias = torch.randn(9, Cout, device=dev); slope = torch.rand(Cout, device=dev)
    o = torch.empty(Bn, H, H, Cout, device=dev, dtype=torch.bfloat16)
    for _ in range(4):
        ctx.frb_debug_conv(L, Bn, x.data_ptr(), None, None, w.data_ptr(), bias.data_ptr(), slope.data_ptr(), o.data_ptr(), 0, None)
    torch.cuda.synchronize()
    t = tr.cpu().view(148, 128)
    t0 = t[:, 0].min()
    names = ["start", "prologue", "weights", "slab0", "slab1", "slab2", "slab3", "mma_done", "epi0", "epi1", "epi_done", "end"]
    print(f"B={Bn} {H}x{H} {Cin}->{Cout}  (us since first CTA start)")
    for cta in (0, 1, 2, 74, 146):
        print(f" cta {cta:3d}: " + "  ".join(f"{n}={(int(t[cta, i]) - int(t0)) / 1e3:7.2f}" if t[cta, i] > 0 else f"{n}=   -   " for i, n in enumerate(names)))
    print(" weights-arrival per leader CTA (us):", " ".join(f"{(int(t[c, 2]) - int(t0)) / 1e3:.0f}" for c in range(0, 148, 2)))
    print(" end per leader CTA (us):", " ".join(f"{(int(t[c, 11]) - int(t0)) / 1e3:.0f}" for c in range(0, 148, 2)))
    for cta in (0, 74):
        row = t[cta]
        print(f" cta {cta} per-tile (us): " + " | ".join(f"acc {(int(row[16+i*4])-int(t0))/1e3:.2f} s0 {(int(row[17+i*4])-int(t0))/1e3:.2f} s1 {(int(row[18+i*4])-int(t0))/1e3:.2f}" for i in range(12) if row[16+i*4] > 0))
    print(f" all CTAs: start max {(int(t[:,0].max())-int(t0))/1e3:.2f}  prologue max {(int(t[:,1].max())-int(t0))/1e3:.2f}  end max {(int(t[:,11].max())-int(t0))/1e3:.2f}")
conv(256, 28, 128, 128); conv(256, 56, 64, 64)
