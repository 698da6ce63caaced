#!/usr/bin/env python
"""Match-only timing on one GPU: P probes x N x 512 gallery, top-5 (BASELINE config 3 per-GPU shape when N = 1M / world)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from facerecognitionpipeline_b200 import _native
ctx = _native.Context(0)
dev = torch.device("cuda", 0)
N = int(os.environ.get("FRB_N", 1_000_000))
g = torch.Generator(device=dev).manual_seed(7)
G = torch.randn((N, 512), generator=g, device=dev); G /= G.norm(dim=1, keepdim=True)
ctx.frb_gallery_upload(G.data_ptr(), N, 0, 1)
st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
for P in [int(a) for a in sys.argv[1:]] or [256, 1024, 4096]:
    probes = G[torch.randint(0, N, (P,), generator=g, device=dev)] + 0.03 * torch.randn((P, 512), generator=g, device=dev)
    probes[::7] = torch.randn((probes[::7].shape[0], 512), generator=g, device=dev)
    if os.environ.get("FRB_PROBES") == "random":      # every probe an impostor (what a shard sees for probes enrolled elsewhere)
        probes = torch.randn((P, 512), generator=g, device=dev)
    sc = torch.empty((P, 5), dtype=torch.float32, device=dev); ix = torch.empty((P, 5), dtype=torch.int64, device=dev)
    ac = torch.empty((P,), dtype=torch.uint8, device=dev)
    f = lambda: ctx.frb_match(probes.data_ptr(), P, 5, 0.35, 1, sc.data_ptr(), ix.data_ptr(), ac.data_ptr(), None, st)
    for _ in range(3): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): f()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(f"P={P:5d} N={N}: {ms:8.3f} ms  {P*1024.0*N/ms/1e9:7.1f} TFLOP/s  gallery pass {N*1024/ms/1e6:7.1f} GB/s-equivalent  flagged {ctx._lib.frb_match_last_flagged(ctx.handle)}", flush=True)
    ms3 = (C.c_float * 3)()
    parts = []
    for _ in range(3):
        ctx.frb_match_profile(probes.data_ptr(), P, 5, 0.35, 1, sc.data_ptr(), ix.data_ptr(), ac.data_ptr(), st, ms3)
        parts.append(list(ms3))
    print("        parts (prepare | filter | finalize + fix-up) ms:", [round(float(x), 4) for x in sorted(parts)[1]], flush=True)
    if os.environ.get("FRB_SHARDED1", "0") == "1" and P <= 4096:
        # the sharded entry point with a world of ONE: probe push + flag wait + match + row push + merge on this GPU alone
        # (what one rank of an N-rank sharded match spends when it never waits for a peer)
        if not getattr(ctx, "_x", False):
            ctx.frb_xchg_create(1, 0, 4096, 8, (C.c_ubyte * 64)())
            ctx._x = True
        fs = lambda: ctx.frb_match_sharded(probes.data_ptr(), 0, P, P, 5, 0.35, 1, sc.data_ptr(), ix.data_ptr(), ac.data_ptr(), st)
        for _ in range(3): fs()
        torch.cuda.synchronize()
        a.record()
        for _ in range(10): fs()
        b.record(); torch.cuda.synchronize()
        print(f"        frb_match_sharded, world 1: {a.elapsed_time(b) / 10:8.3f} ms", flush=True)
