#!/usr/bin/env python
"""Warp + normalise (HBM-bound; 225 792 algorithmic bytes per face at source scale 2) and the config-4 stream
(warp -> IR-101 embed -> match at batch 1024) on one GPU.  Device-resident frames, CUDA-event timing.
Every face has its own 256x256 source frame (no frame re-use: the source really comes from HBM)."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from facerecognitionpipeline_b200 import _native, weights
from facerecognitionpipeline_b200.face_recognition import similarity_template, estimate_matrix

ctx = _native.Context(0)
dev = torch.device("cuda", 0)
B, S, H = int(os.environ.get("FRB_B", 1024)), 112, 256
rng = np.random.default_rng(3)
tpl = similarity_template(S)
jobs = (_native.WarpJob * B)()
area = 0.0
for i in range(B):
    ang, sc = np.deg2rad(rng.uniform(-20, 20)), rng.uniform(1.5, 2.2)
    R = np.array([[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]]) * sc
    lm = ((tpl - S / 2) @ R.T + np.array([128 + rng.uniform(-8, 8), 128 + rng.uniform(-8, 8)]) + rng.normal(0, 0.5, (5, 2))).astype(np.float32)
    M = estimate_matrix(lm, tpl)
    jobs[i].src_off, jobs[i].H, jobs[i].W, jobs[i].pitch = i * H * H * 3, H, H, H * 3
    for j, v in enumerate(np.asarray(M, np.float64).reshape(6)):
        jobs[i].M[j] = float(v)
    area += (S * sc) ** 2          # source footprint of the crop in pixels
frames = torch.randint(0, 256, (B, H, H, 3), dtype=torch.uint8, device=dev)
x = torch.empty((B, S, S, 3), dtype=torch.bfloat16, device=dev)
st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timed(fn, reps=20, flush_l2=True):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        if flush_l2: flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / reps

peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) if os.path.exists(
    os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
ms = timed(lambda: ctx.frb_warp_normalize(frames.data_ptr(), jobs, B, S, None, x.data_ptr(), st))
alg = area * 3 + B * S * S * 3 * 2            # source footprint bytes + bf16 NHWC output
print(json.dumps(dict(kernel="warp_normalize_kernel", faces=B, ms=ms, faces_per_s=B / ms * 1e3, algorithmic_bytes_per_face=alg / B,
                      achieved_gbs=alg / ms / 1e6, peak_gbs=peaks["hbm_gbs"], frac=alg / ms / 1e6 / peaks["hbm_gbs"],
                      note="L2 flushed between repetitions; includes the host->device copy of the job table (72 B per face)")), flush=True)
crops = torch.randint(0, 256, (B, S, S, 3), dtype=torch.uint8, device=dev)
ms = timed(lambda: ctx.frb_preprocess_u8(crops.data_ptr(), B, S, x.data_ptr(), 0, st))
alg = B * S * S * 3 * 3
print(json.dumps(dict(kernel="preprocess_u8_kernel", faces=B, ms=ms, faces_per_s=B / ms * 1e3, algorithmic_bytes_per_face=alg / B,
                      achieved_gbs=alg / ms / 1e6, peak_gbs=peaks["hbm_gbs"], frac=alg / ms / 1e6 / peaks["hbm_gbs"])), flush=True)
if os.environ.get("FRB_STREAM", "1") == "1":
    prog = weights.build_program(weights.random_init_state_dict("ir_101", "adaface", seed=0), "ir_101", "adaface")
    prog.load_into(ctx)
    N = 1_000_000
    g = torch.Generator(device=dev).manual_seed(1)
    G = torch.randn((N, 512), generator=g, device=dev); G /= G.norm(dim=1, keepdim=True)
    ctx.frb_gallery_upload(G.data_ptr(), N, 0, 1); del G
    emb = torch.empty((B, 512), dtype=torch.float32, device=dev)
    sc = torch.empty((B, 5), dtype=torch.float32, device=dev); ix = torch.empty((B, 5), dtype=torch.int64, device=dev)
    ac = torch.empty((B,), dtype=torch.uint8, device=dev)
    flags = _native.FRB_EMBED_L2 | _native.FRB_EMBED_RENORM
    def step():
        ctx.frb_warp_normalize(frames.data_ptr(), jobs, B, S, None, x.data_ptr(), st)
        ctx.frb_embed(x.data_ptr(), B, flags, emb.data_ptr(), None, None, st)
        ctx.frb_match(emb.data_ptr(), B, 5, 0.35, 1, sc.data_ptr(), ix.data_ptr(), ac.data_ptr(), None, st)
    ms = timed(step, reps=10, flush_l2=False)
    print(json.dumps(dict(stream="config 4: warp -> IR-101 embed -> match vs 1M, batch %d" % B, ms_per_batch=ms, faces_per_s=B / ms * 1e3)), flush=True)
