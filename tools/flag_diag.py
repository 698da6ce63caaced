#!/usr/bin/env python
"""Why does match_finalize_kernel flag a row?  Emulates the proof of simple_kernels.cuh (match_finalize_kernel) with
torch on the GPU - static per-slice top-8 lists (no shared floors), the 41st best candidate, the exact 5th score -
and prints, for the rows whose margin is smallest, which of the two bounds is the binding one.
    FRB_N=125000 python tools/flag_diag.py 4096
"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from facerecognitionpipeline_b200 import _native
ctx = _native.Context(0)
dev = torch.device("cuda", 0)
N = int(os.environ.get("FRB_N", 125_000))
P = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
K, KCAND, R = 5, 8, 40
g = torch.Generator(device=dev).manual_seed(7)
G = torch.randn((N, 512), generator=g, device=dev); G /= G.norm(dim=1, keepdim=True)
ctx.frb_gallery_upload(G.data_ptr(), N, 0, 1)
st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
probes = G[torch.randint(0, N, (P,), generator=g, device=dev)] + 0.03 * torch.randn((P, 512), generator=g, device=dev)
probes[::7] = torch.randn((probes[::7].shape[0], 512), generator=g, device=dev)
sc = torch.empty((P, K), dtype=torch.float32, device=dev); ix = torch.empty((P, K), dtype=torch.int64, device=dev)
ac = torch.empty((P,), dtype=torch.uint8, device=dev)
for _ in range(3):
    ctx.frb_match(probes.data_ptr(), P, K, 0.35, 1, sc.data_ptr(), ix.data_ptr(), ac.data_ptr(), None, st)
    torch.cuda.synchronize()
    print("frb_match flagged rows:", ctx._lib.frb_match_last_flagged(ctx.handle), flush=True)

q = probes / probes.norm(dim=1, keepdim=True)
qb, Gb = q.bfloat16().float(), G.bfloat16().float()
# slice geometry of match_core (pair mode)
units, p_tiles, g_tiles = 74, (P + 255) // 256, (N + 255) // 256
smin = max(1, -(-units // p_tiles)); smax = max(1, min(g_tiles, 2048 // KCAND, smin * 8))
best = None
for sl in range(min(smin, smax), smax + 1):
    tps = -(-g_tiles // sl); real = -(-g_tiles // tps)
    cost = -(-(p_tiles * real) // units) * (tps + 2)
    if best is None or cost < best[0]: best = (cost, real, tps)
_, slices, tps = best
rows_per_slice = tps * 256
print(f"N {N} P {P}: {slices} slices of {tps} tiles ({rows_per_slice} rows)")
torch.backends.cuda.matmul.allow_tf32 = False
approx = qb @ Gb.T                                   # [P, N] fp32 accumulate of bf16 products
exact = (q.double() @ G.double().T)
s5 = exact.topk(K, dim=1).values[:, K - 1]
pad = slices * rows_per_slice - N
ap = torch.nn.functional.pad(approx, (0, pad), value=float("-inf")).view(P, slices, rows_per_slice)
top8 = ap.topk(KCAND, dim=2).values                 # per-slice lists
excl = top8[:, :, KCAND - 1].max(dim=1).values      # bound on everything a slice dropped
cand = top8.reshape(P, -1).sort(dim=1, descending=True).values
s41 = cand[:, R]
gmax = G.norm(dim=1).max(); gerr = (G - Gb).norm(dim=1).max(); qerr = (q - qb).norm(dim=1); qbn = qb.norm(dim=1)
eps_r1 = (2 ** -7 + 2 ** -12) * 1.0001 + 1e-6          # round 1: element-wise worst case
eps = (qerr * gmax + qbn * gerr + 2 ** -12 * qbn * (gmax + gerr)) * 1.0001 + 1e-6     # match_finalize_kernel now
print(f"eps: round 1 {eps_r1:.5f}, now mean {eps.mean().item():.5f} max {eps.max().item():.5f}  (||q-qb|| mean {qerr.mean().item():.5f}, max||g-gb|| {gerr.item():.5f})")
bound = torch.maximum(excl, s41)
margin = s5.float() - (bound + eps)
flag = margin <= 0
print(f"with the round-1 eps: {int((s5.float() - (bound + eps_r1) <= 0).sum())} flagged")
print(f"emulated (no shared floors): {int(flag.sum())} flagged; binding bound = slice-8th in {int((excl >= s41).sum())} rows, 41st candidate in {int((excl < s41).sum())}")
order = margin.argsort()[:12]
for r in order.tolist():
    print(f"  row {r:5d} (random probe: {r % 7 == 0}): s1 {exact[r].max().item():.4f} s5 {s5[r].item():.4f}  slice-8th max {excl[r].item():.4f}  41st {s41[r].item():.4f}  eps {eps[r].item():.4f}  margin {margin[r].item():+.4f}")
qs = torch.tensor([0.001, 0.01, 0.1, 0.5], device=dev)
print("margin quantiles (0.1 %, 1 %, 10 %, 50 %):", [round(v, 4) for v in margin.quantile(qs).tolist()])
print("s5 - s41 quantiles:", [round(v, 4) for v in (s5.float() - s41).quantile(qs).tolist()], " s5 - excl:", [round(v, 4) for v in (s5.float() - excl).quantile(qs).tolist()])
err = (exact - approx.double()).abs().max(dim=1).values.float()   # torch's bf16-product sums stand in for the TMEM accumulation
print(f"observed max |exact - approx| over all {P} x {N} pairs: {err.max().item():.5f}; largest observed/eps ratio {float((err / eps).max()):.3f} (must stay below 1)")
