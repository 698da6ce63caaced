#!/usr/bin/env python
"""Bring-up ladder for the sm_100a kernels: each rung runs in its own process under a timeout so a
faulting rung cannot take the later ones down.  Results go to gpurun_out/ladder_<rung>.json.

    python tools/gpu_ladder.py            # all rungs
    python tools/gpu_ladder.py gemm       # one rung (inside the child process)
"""
import json
import os
import subprocess
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")
RUNGS = ["gemm", "im2col", "conv", "match", "conv_big"]


def _ctx():
    import torch  # noqa
    from facerecognitionpipeline_b200 import _native
    return _native.Context(0)


def bf16_round(x):
    import torch
    return x.to(torch.bfloat16).to(torch.float32)


def rung_gemm():
    import torch
    ctx = _ctx()
    res = []
    g = torch.Generator(device="cpu").manual_seed(0)
    for (M, N, K, splits) in [(128, 256, 64, 1), (128, 256, 512, 1), (128, 64, 128, 1), (128, 128, 128, 1),
                               (300, 512, 1024, 1), (256, 512, 25088, 37), (4096, 1024, 512, 1)]:
        A = (torch.randn(M, K, generator=g) * 0.5).to(torch.bfloat16).cuda()
        B = (torch.randn(N, K, generator=g) * 0.5).to(torch.bfloat16).cuda()
        import facerecognitionpipeline_b200._native as nat
        # realised split count mirrors choose_splits in api.cu
        nkb = K // 64
        per = -(-nkb // max(1, min(splits, nkb)))
        real = -(-nkb // per)
        Cc = torch.zeros(real, M, N, dtype=torch.float32, device="cuda")
        ctx.frb_debug_gemm(A.data_ptr(), B.data_ptr(), M, N, K, real, Cc.data_ptr(), None)
        torch.cuda.synchronize()
        got = Cc.sum(0)
        ref = A.float() @ B.float().t()
        err = (got - ref).abs().max().item()
        rel = err / ref.abs().max().item()
        res.append(dict(M=M, N=N, K=K, splits=real, max_abs_err=err, rel=rel, ok=bool(rel < 2e-3)))
        print(res[-1], flush=True)
    return res


def rung_im2col():
    import torch
    ctx = _ctx()
    res = []
    for (B, H, W, Cc, ks, st, pd, m0, c0, r, s) in [
        (2, 14, 14, 64, 3, 1, 1, 0, 0, 0, 0),
        (2, 14, 14, 64, 3, 1, 1, 0, 0, 1, 1),
        (2, 14, 14, 64, 3, 1, 1, 128, 0, 2, 2),
        (2, 14, 14, 128, 3, 1, 1, 128, 64, 0, 2),
        (3, 28, 28, 64, 3, 2, 1, 128, 0, 1, 0),
        (3, 28, 28, 64, 1, 2, 0, 128, 0, 0, 0),
        (1, 112, 112, 64, 3, 2, 1, 3072, 0, 2, 1),
    ]:
        x = torch.zeros(B, H, W, Cc)
        n_i, h_i, w_i = torch.meshgrid(torch.arange(B), torch.arange(H), torch.arange(W), indexing="ij")
        x[..., 0::4] = n_i[..., None].float() + 1
        x[..., 1::4] = h_i[..., None].float() + 1
        x[..., 2::4] = w_i[..., None].float() + 1
        x[..., 3::4] = torch.arange(Cc // 4).float()[None, None, None, :]
        xd = x.to(torch.bfloat16).cuda()
        out = torch.zeros(128 * 64, dtype=torch.bfloat16, device="cuda")
        ctx.frb_debug_im2col(xd.data_ptr(), B, H, W, Cc, ks, st, pd, m0, c0, r, s, out.data_ptr(), None)
        torch.cuda.synchronize()
        raw = out.float().cpu().view(128, 8, 8)  # row, 16B chunk, 8 elems
        # undo the 128B swizzle: physical chunk = logical chunk ^ (row & 7)
        tile = torch.zeros(128, 8, 8)
        for row in range(128):
            for ch in range(8):
                tile[row, ch] = raw[row, ch ^ (row & 7)]
        tile = tile.view(128, 64)
        P = (H + 2 * pd - ks) // st + 1
        Q = (W + 2 * pd - ks) // st + 1
        exp = torch.zeros(128, 64)
        for row in range(128):
            m = m0 + row
            img, rem = divmod(m, P * Q)
            pp, qq = divmod(rem, Q)
            iy, ix = pp * st - pd + r, qq * st - pd + s
            if img < B and 0 <= iy < H and 0 <= ix < W:
                exp[row] = x[img, iy, ix, c0:c0 + 64]
        bad = (tile != exp).any(dim=1)
        nbad = int(bad.sum())
        info = dict(B=B, H=H, W=W, C=Cc, ks=ks, st=st, pd=pd, m0=m0, c0=c0, r=r, s=s, bad_rows=nbad, ok=nbad == 0)
        if nbad:
            rows = [i for i in range(128) if bad[i]][:12]
            info["examples"] = [dict(row=i, got_nhw=[tile[i, 0].item(), tile[i, 1].item(), tile[i, 2].item(), tile[i, 3].item()],
                                     exp_nhw=[exp[i, 0].item(), exp[i, 1].item(), exp[i, 2].item(), exp[i, 3].item()]) for i in rows]
        res.append(info)
        print(info, flush=True)
    return res


def _conv_case(ctx, B, H, W, Cin, Cout, ks, st, cases, prelu, res_mode, sc_cin, seed):
    """res_mode: 0 none, 1 identity residual (same size), 2 strided identity (MaxPool(1,2))"""
    import torch
    from facerecognitionpipeline_b200._native import LayerDesc
    g = torch.Generator(device="cpu").manual_seed(seed)
    pd = 1 if ks == 3 else 0
    P = (H + 2 * pd - ks) // st + 1
    Q = (W + 2 * pd - ks) // st + 1
    x = torch.randn(B, H, W, Cin, generator=g).to(torch.bfloat16).cuda()
    ktot = ks * ks * Cin + sc_cin
    w = (torch.randn(Cout, ktot, generator=g) / (ktot ** 0.5)).to(torch.bfloat16).cuda()
    bias = torch.randn(cases, Cout, generator=g).float().cuda()
    slope = (torch.rand(Cout, generator=g) * 0.3 + 0.1).float().cuda()
    L = LayerDesc()
    L.op = 1; L.cin = Cin; L.cout = Cout; L.hin = H; L.win = W; L.ksize = ks; L.stride = st; L.pad = pd
    L.in_buf = 0; L.out_buf = 1; L.sc_buf = -1; L.res_buf = -1; L.bias_cases = cases; L.has_prelu = 1 if prelu else 0
    sc = None
    if sc_cin:
        SH, SW = P * 2, Q * 2
        sc = torch.randn(B, SH, SW, sc_cin, generator=g).to(torch.bfloat16).cuda()
        L.sc_buf = 2; L.sc_cin = sc_cin; L.sc_hin = SH; L.sc_win = SW; L.sc_stride = 2
    resid = None
    if res_mode == 1:
        resid = torch.randn(B, P, Q, Cout, generator=g).to(torch.bfloat16).cuda()
        L.res_buf = 3; L.res_h = P; L.res_w = Q; L.res_stride = 1
    elif res_mode == 2:
        resid = torch.randn(B, 2 * P, 2 * Q, Cout, generator=g).to(torch.bfloat16).cuda()
        L.res_buf = 3; L.res_h = 2 * P; L.res_w = 2 * Q; L.res_stride = 2
    outs = []
    for use_ref in (1, 0):
        o = torch.full((B, P, Q, Cout), float("nan"), dtype=torch.bfloat16, device="cuda")
        ctx.frb_debug_conv(L, B, x.data_ptr(), sc.data_ptr() if sc is not None else None,
                           resid.data_ptr() if resid is not None else None, w.data_ptr(), bias.data_ptr(),
                           slope.data_ptr(), o.data_ptr(), use_ref, None)
        torch.cuda.synchronize()
        outs.append(o.float())
    # independent torch check of the reference kernel
    xf = x.float().permute(0, 3, 1, 2)
    wf = w.float()[:, :ks * ks * Cin].view(Cout, ks, ks, Cin).permute(0, 3, 1, 2)
    y = torch.nn.functional.conv2d(xf, wf, stride=st, padding=pd)
    if sc is not None:
        wsc = w.float()[:, ks * ks * Cin:].view(Cout, sc_cin, 1, 1)
        y = y + torch.nn.functional.conv2d(sc.float().permute(0, 3, 1, 2), wsc, stride=2)
    y = y.permute(0, 2, 3, 1)
    if cases == 9:
        rc = torch.ones(P, dtype=torch.long); rc[0] = 0; rc[-1] = 2
        cc = torch.ones(Q, dtype=torch.long); cc[0] = 0; cc[-1] = 2
        case = (rc[:, None] * 3 + cc[None, :]).cuda()
        y = y + bias[case][None]
    else:
        y = y + bias[0]
    if prelu:
        y = torch.where(y > 0, y, y * slope)
    if res_mode == 1:
        y = y + resid.float()
    elif res_mode == 2:
        y = y + resid.float()[:, ::2, ::2]
    scale = y.abs().max().item()
    e_ref = (outs[0] - y).abs().max().item() / scale
    e_tc = (outs[1] - y).abs().max().item() / scale
    e_tc_ref = (outs[1] - outs[0]).abs().max().item() / scale
    nan = bool(torch.isnan(outs[1]).any())
    return dict(B=B, H=H, W=W, Cin=Cin, Cout=Cout, ks=ks, st=st, cases=cases, prelu=prelu, res_mode=res_mode,
                sc_cin=sc_cin, err_ref_vs_torch=e_ref, err_tc_vs_torch=e_tc, err_tc_vs_ref=e_tc_ref, nan=nan,
                ok=bool(e_tc < 1.5e-2 and e_ref < 1.5e-2 and not nan))


def rung_conv():
    ctx = _ctx()
    res = []
    cases = [
        # B, H, W, Cin, Cout, ks, st, cases, prelu, res_mode, sc_cin
        (2, 14, 14, 64, 64, 3, 1, 1, False, 0, 0),
        (2, 14, 14, 64, 64, 3, 1, 9, True, 0, 0),
        (3, 14, 14, 256, 256, 3, 1, 9, True, 0, 0),
        (3, 14, 14, 256, 256, 3, 1, 1, False, 1, 0),
        (2, 28, 28, 128, 128, 3, 2, 1, False, 0, 0),
        (2, 28, 28, 128, 256, 3, 2, 1, False, 0, 128),
        (2, 112, 112, 64, 64, 3, 2, 1, False, 2, 0),
        (5, 7, 7, 512, 512, 3, 1, 9, True, 0, 0),
        (5, 14, 14, 512, 512, 3, 2, 1, False, 0, 256),
        (4, 56, 56, 64, 128, 3, 1, 9, True, 0, 0),
        # activation-slab kernel shapes (3x3 stride 1 at W = 112 / 56 / 28), incl. odd tile counts
        (2, 56, 56, 64, 64, 3, 1, 1, False, 1, 0),
        (3, 56, 56, 64, 64, 3, 1, 9, True, 0, 0),
        (3, 28, 28, 128, 128, 3, 1, 9, True, 0, 0),
        (3, 28, 28, 128, 128, 3, 1, 1, False, 1, 0),
        (5, 28, 28, 128, 256, 3, 1, 9, True, 0, 0),
        (1, 112, 112, 64, 64, 3, 1, 9, True, 0, 0),
        (37, 28, 28, 128, 128, 3, 1, 9, True, 0, 0),
    ]
    for i, c in enumerate(cases):
        r = _conv_case(ctx, *c, seed=i)
        res.append(r)
        print(r, flush=True)
    return res


def rung_conv_big():
    """Full-batch shapes whose last round of tiles is partial: exercises FRB_TAIL_SPLIT (split-K tail) and FRB_QUAD."""
    ctx = _ctx()
    res = []
    cases = [
        (256, 14, 14, 256, 256, 3, 1, 9, True, 0, 0),     # 196 tiles on 74 pairs: 48 tail tiles x 3
        (256, 14, 14, 256, 256, 3, 1, 1, False, 1, 0),
        (200, 14, 14, 256, 256, 3, 1, 9, True, 1, 0),     # 154 tiles: 6 tail tiles
        (256, 7, 7, 512, 512, 3, 1, 9, True, 0, 0),       # two N tiles, 24 tail tiles
        (64, 28, 28, 128, 256, 3, 1, 9, True, 0, 0),      # 196 tiles, K = 18 blocks
        (250, 14, 14, 256, 512, 3, 1, 9, True, 0, 0),     # M not a multiple of 256 + two N tiles
    ]
    for i, c in enumerate(cases):
        r = _conv_case(ctx, *c, seed=100 + i)
        res.append(r)
        print(r, flush=True)
    return res


def rung_match():
    import numpy as np
    import torch
    ctx = _ctx()
    res = []
    rng = np.random.default_rng(0)
    for (N, P, k) in [(23, 5, 3), (100, 32, 5), (5000, 64, 5), (100000, 300, 5), (1000000, 256, 5)]:
        G = rng.standard_normal((N, 512), dtype=np.float32)
        G /= np.linalg.norm(G, axis=1, keepdims=True)
        sel = rng.integers(0, N, size=P)
        Pm = G[sel] + 0.05 * rng.standard_normal((P, 512), dtype=np.float32)
        Pm[P // 2:] = rng.standard_normal((P - P // 2, 512), dtype=np.float32)
        Pm = (Pm / (np.linalg.norm(Pm, axis=1, keepdims=True) + 1e-8)).astype(np.float32)
        ctx.frb_gallery_upload(G.ctypes.data, N, 0, 0)
        sc = np.zeros((P, k), np.float32)
        ix = np.zeros((P, k), np.int64)
        ac = np.zeros((P,), np.uint8)
        t0 = time.time()
        ctx.frb_match_host(Pm.ctypes.data, P, k, 0.35, 0, sc.ctypes.data, ix.ctypes.data, ac.ctypes.data)
        dt = time.time() - t0
        S = G.astype(np.float64) @ Pm.astype(np.float64).T  # [N, P]
        ok_idx = True
        worst = 0.0
        for p_ in range(P):
            s = S[:, p_]
            order = np.lexsort((np.arange(N), -s))[:k]
            kk = min(k, N)
            if not np.array_equal(order[:kk], ix[p_, :kk]):
                ok_idx = False
            worst = max(worst, float(np.abs(s[order[:kk]] - sc[p_, :kk]).max()))
        flagged = int(ctx.frb_match_last_flagged())
        info = dict(N=N, P=P, k=k, idx_exact=ok_idx, max_score_err=worst, flagged=flagged, secs=dt,
                    ok=bool(ok_idx and worst < 1e-6))
        res.append(info)
        print(info, flush=True)
    return res


def main():
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] in RUNGS:
        name = sys.argv[1]
        try:
            r = globals()["rung_" + name]()
            status = "ok" if all(x.get("ok") for x in r) else "mismatch"
        except Exception:
            r = traceback.format_exc()
            status = "exception"
            print(r, flush=True)
        with open(os.path.join(OUT, f"ladder_{name}.json"), "w") as f:
            json.dump(dict(status=status, results=r), f, indent=1)
        sys.exit(0 if status == "ok" else 1)
    rungs = sys.argv[1:] or RUNGS
    summary = {}
    for name in rungs:
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), name], timeout=240, capture_output=True, text=True)
            rc, tail = p.returncode, (p.stdout + p.stderr)[-6000:]
        except subprocess.TimeoutExpired as e:
            rc, tail = -9, "TIMEOUT " + str((e.stdout or b"")[-3000:])
        summary[name] = dict(rc=rc, secs=round(time.time() - t0, 1))
        with open(os.path.join(OUT, f"ladder_{name}.log"), "w") as f:
            f.write(tail)
        print(f"== {name}: rc={rc} in {summary[name]['secs']}s\n{tail[-3000:]}", flush=True)
    with open(os.path.join(OUT, "ladder_summary.json"), "w") as f:
        json.dump(summary, f, indent=1)


if __name__ == "__main__":
    main()
