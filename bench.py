#!/usr/bin/env python
"""bench.py — headline benchmark of the embed-and-match hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one batch of synthetic aligned 112x112 RGB crops per GPU -> IR-101 (AdaFace layout,
random-init weights) -> L2-norm -> cosine match against a 1M x 512 synthetic gallery (top-5 +
threshold).  `value` is timed with the crops already resident in HBM; `e2e` goes through the C ABI
entry point with pinned HOST buffers (H2D of the crops and D2H of the results inside the timed
region).  Multi-GPU = probe data-parallel (gallery + weights replicated, no data-path collective),
weak scaling.  Rank 0 prints ONE JSON line.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "aligned faces/sec embedded+matched (IR-101, 1M gallery)"
# dram__bytes_read.sum + dram__bytes_write.sum of the persistent 66-layer run at batch 256 (ncu --set full of one launch, profiles/r02_summary.md /
# r02a_raw_multi.csv; round 1 measured 305.98 + 1145.59 MB)
RUN_TRAFFIC = 293.33e6 + 1198.14e6
UNIT = "faces/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=float(d["hbm_gbs"]), tf_burst=float(d["bf16_tflops"]),
                    tf_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler:
    """Samples SM clock / throttle reasons during the timed region: NVML polled every 5 ms from a thread (the timed
    region of a default run is only ~0.15 s, too short for `nvidia-smi -lms`), nvidia-smi as the fallback."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.sm, self.power, self.mx, self.reasons, self.src = gpu_index, [], [], None, set(), None
        self._stop = threading.Event()
        self.t = None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.gpu < len(ids) and ids[self.gpu].isdigit():
                return int(ids[self.gpu])
        return self.gpu

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                    "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                    "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}

            def poll():
                while not self._stop.is_set():
                    try:
                        self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                        for name, bit in bits.items():
                            if r & bit:
                                self.reasons.add(name)
                        self.power.append(nv.nvmlDeviceGetPowerUsage(h) / 1e3)
                    except Exception:
                        pass
                    time.sleep(0.005)

            self.src = "nvml"
            self.t = threading.Thread(target=poll, daemon=True)
            self.t.start()
        except Exception:
            self.src = None

    def _smi_once(self):
        try:
            out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                  str(self._physical_index())], capture_output=True, text=True, timeout=10).stdout
            f = [x.strip() for x in out.strip().splitlines()[0].split(",")]
            self.sm.append(float(f[1]))
            self.mx = float(f[2])
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    self.reasons.add(name)
            self.src = "nvidia-smi"
        except Exception:
            pass

    def stop(self):
        """Call while the GPU is still busy with the last timed step (before the closing synchronize)."""
        if self.t is None or not self.sm:
            self._smi_once()
        self._stop.set()
        if self.t is not None:
            self.t.join(timeout=1)
        if not self.sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["clock query unavailable"], samples=0)
        return dict(sm_mhz=float(np.median(self.sm)), sm_max_mhz=self.mx, reasons=sorted(self.reasons),
                    samples=len(self.sm), power_w_max=max(self.power) if self.power else None, source=self.src)


# ------------------------------------------------------------------------------------------------
class CpuReference:
    """The reference's CPU path on this box's host cores: torch-eager fp32 IR-101 in batches of 32
    (face_embedder.py:137-182, via the oracle restatement) + numpy matching exactly as
    GalleryManager.search does per probe with the matrix cached (gallery_manager.py:195-197).
    Model and gallery are built once; `sample` times one bounded sample and returns seconds per face / per probe."""

    def __init__(self, gallery_rows, seed=0, threads=None):
        import torch
        from oracle import backbone, embedder
        self.threads = threads or os.cpu_count()
        torch.set_num_threads(self.threads)
        try:   # torchrun exports OMP_NUM_THREADS=1: give numpy's BLAS (the matching half) all host threads back
            import threadpoolctl
            self._blas_limit = threadpoolctl.threadpool_limits(limits=self.threads)
        except Exception:
            self._blas_limit = None
        self.rng = np.random.default_rng(seed)
        sd = backbone.random_state_dict("ir_101", "adaface", seed, calibrate=False)
        self.emb = embedder.OracleEmbedder("ir_101", "adaface", state_dict=sd)
        G = self.rng.standard_normal((gallery_rows, 512), dtype=np.float32)
        G /= np.linalg.norm(G, axis=1, keepdims=True)
        self.G = G
        self.emb.extract_embeddings_batch([self.rng.integers(0, 256, (112, 112, 3), dtype=np.uint8) for _ in range(2)])

    def sample(self, n_faces, n_probes):
        crops = [self.rng.integers(0, 256, (112, 112, 3), dtype=np.uint8) for _ in range(n_faces)]
        t0 = time.perf_counter()
        E = self.emb.extract_embeddings_batch(crops, normalize=True, batch_size=32)
        t_embed = (time.perf_counter() - t0) / n_faces
        t0 = time.perf_counter()
        for i in range(n_probes):
            q = E[i % len(E)]
            q = q / (np.linalg.norm(q) + 1e-8)
            s = np.dot(self.G, q)
            _ = np.argsort(s)[::-1][:5]
        t_match = (time.perf_counter() - t0) / n_probes
        return t_embed, t_match

    def sample_as_is(self, n_probes=2):
        """GalleryManager.search exactly as the reference runs it (gallery_manager.py:177-205): EVERY call rebuilds the
        matrix with np.vstack over the students' templates before the dot product and the full argsort.  Seconds per
        probe; a few probes only (SURVEY §8d 'as-is')."""
        students = {i: self.G[i] for i in range(len(self.G))}      # sid -> template_embedding, dict insertion order
        q = self.G[0]
        t0 = time.perf_counter()
        for _ in range(n_probes):
            ids = list(students.keys())
            E = np.vstack([students[sid] for sid in ids])
            qq = q / (np.linalg.norm(q) + 1e-8)
            s = np.dot(E, qq)
            _ = np.argsort(s)[::-1][:5]
        return (time.perf_counter() - t0) / n_probes


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (oracle port; the
    reference modules themselves cannot be imported: `import net` fails)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    gallery_rows = args.gallery
    ref = CpuReference(gallery_rows)
    n_faces, n_probes = 32, 4
    per_step = []
    t_begin = time.perf_counter()
    for i in range(args.warmup + args.steps):
        te, tm = ref.sample(n_faces, n_probes)
        if i >= args.warmup:
            per_step.append(te + tm)
        if i == 0 and (time.perf_counter() - t_begin) * (args.warmup + args.steps) > 240:
            n_faces, n_probes = 8, 2              # slow host: keep the whole run within a few minutes
    sec_per_face = float(np.mean(per_step))
    value = 1.0 / sec_per_face
    threads = ref.threads
    sample = (f"per step: {n_faces} faces torch-eager fp32 IR-101 (batch 32) + {n_probes} probes numpy dot+argsort vs "
              f"{gallery_rows} x 512 f32 (matrix cached); faces/s = 1 / (embed s/face + match s/probe)")
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=sec_per_face * 1e3 * n_faces, higher_is_better=True, scaling="weak", vs_baseline=None,
                ms_per_step_note=f"extrapolated: {n_faces} x (embed s/face + match s/probe); a step actually times {n_faces} faces "
                                 f"embedded but only {n_probes} probes matched (bounded sample of the same workload)",
                dtype="f32", data="synthetic", impl="reference",
                config=dict(workload=f"IR-101 embed + top-5 match vs {gallery_rows}-identity gallery, CPU reference path",
                            batch_per_gpu=n_faces, gallery_rows=gallery_rows),
                cpu_baseline=dict(value=value, unit=UNIT, cores=threads, kind="port", sample=sample),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from facerecognitionpipeline_b200 import _native, weights

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    ctx = _native.Context(local_rank)
    B, K, W, N, topk, thr = args.batch, args.steps, args.warmup, args.gallery, 5, 0.35

    # weights: random-init IR-101 (no checkpoints exist offline), folded + packed on the host
    sd = weights.random_init_state_dict("ir_101", "adaface", seed=0)
    prog = weights.build_program(sd, "ir_101", "adaface")
    prog.load_into(ctx)
    flops_face = ctx._lib.frb_backbone_flops_per_face(ctx.handle)
    flags = _native.FRB_EMBED_L2 | _native.FRB_EMBED_RENORM

    # gallery: N x 512 unit rows generated on the device (never exists on the host)
    g = torch.Generator(device=dev).manual_seed(1234)
    G = torch.empty((N, 512), dtype=torch.float32, device=dev)
    for s0 in range(0, N, 1 << 18):
        blk = torch.randn((min(1 << 18, N - s0), 512), generator=g, device=dev)
        G[s0:s0 + blk.shape[0]] = blk / blk.norm(dim=1, keepdim=True)
    ctx.frb_gallery_upload(G.data_ptr(), N, 0, 1)
    if world == 1:
        del G          # (world > 1: the identity-sharded section below re-uploads this rank's rows of the same gallery)

    # inputs: NBUF distinct crop batches (rotated so no step re-reads a warm input)
    NBUF = 16
    rng = np.random.default_rng(100 + rank)
    host_crops = torch.from_numpy(rng.integers(0, 256, (NBUF, B, 112, 112, 3), dtype=np.uint8)).pin_memory()
    dev_crops = host_crops.to(dev)
    x_bf16 = torch.empty((B, 112, 112, 3), dtype=torch.bfloat16, device=dev)
    emb = torch.empty((B, 512), dtype=torch.float32, device=dev)
    sc = torch.empty((B, topk), dtype=torch.float32, device=dev)
    ix = torch.empty((B, topk), dtype=torch.int64, device=dev)
    ac = torch.empty((B,), dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev)
    st = C.c_void_p(stream.cuda_stream)

    def step_device(i):
        ctx.frb_preprocess_u8(dev_crops[i % NBUF].data_ptr(), B, 112, x_bf16.data_ptr(), 0, st)
        ctx.frb_embed(x_bf16.data_ptr(), B, flags, emb.data_ptr(), None, None, st)
        ctx.frb_match(emb.data_ptr(), B, topk, thr, 1, sc.data_ptr(), ix.data_ptr(), ac.data_ptr(), None, st)

    h_sc = torch.empty((B, topk), dtype=torch.float32).pin_memory()
    h_ix = torch.empty((B, topk), dtype=torch.int64).pin_memory()
    h_ac = torch.empty((B,), dtype=torch.uint8).pin_memory()

    def step_host(i, more=True):
        # a stream of batches through the public host-buffer API: the NEXT batch's crops are handed to frb_prefetch_host
        # first (copy stream), then this batch is embedded + matched; both copies are inside the timed region
        if more:
            ctx.frb_prefetch_host(host_crops[(i + 1) % NBUF].data_ptr(), B, 112)
        ctx.frb_embed_match_host(host_crops[i % NBUF].data_ptr(), B, 112, flags, topk, thr, None, h_sc.data_ptr(),
                                 h_ix.data_ptr(), h_ac.data_ptr())

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing (value)
    for i in range(W):
        step_device(i)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = ctx.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    t_embed_ms = 0.0
    e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    seg = []
    e_start.record(stream)
    for i in range(K):
        a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        ctx.frb_preprocess_u8(dev_crops[(W + i) % NBUF].data_ptr(), B, 112, x_bf16.data_ptr(), 0, st)
        a.record(stream)
        ctx.frb_embed(x_bf16.data_ptr(), B, flags, emb.data_ptr(), None, None, st)
        b.record(stream)
        ctx.frb_match(emb.data_ptr(), B, topk, thr, 1, sc.data_ptr(), ix.data_ptr(), ac.data_ptr(), None, st)
        c.record(stream)
        seg.append((a, b, c))
    e_end.record(stream)
    e_end.synchronize()
    clocks = sampler.stop()
    barrier()
    launches = ctx.launch_count() - launches0
    ms_total = max_over_ranks(e_start.elapsed_time(e_end))
    embed_ms = float(np.mean([a.elapsed_time(b) for a, b, c in seg]))
    match_ms = float(np.mean([b.elapsed_time(c) for a, b, c in seg]))
    value = world * B * K / (ms_total / 1e3)

    # ---- end-to-end through the host-buffer C ABI (e2e)
    for i in range(max(W, 3)):
        step_host(i)
    barrier()
    t0 = time.perf_counter()
    ctx.frb_prefetch_host(host_crops[W % NBUF].data_ptr(), B, 112)   # batch 0 of the stream: its copy is not overlapped
    for i in range(K):
        step_host(W + i, more=(i + 1 < K))
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * B * K / e2e_s
    h2d = B * 112 * 112 * 3
    d2h = B * topk * 4 + B * topk * 8 + B

    peaks = load_peaks()

    def timed(fn, reps, warm=3):
        """CUDA-event time per call (ms) on the launch stream, barrier on both sides, max over ranks."""
        for _ in range(warm):
            fn()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(reps):
            fn()
        b.record(stream)
        b.synchronize()
        barrier()
        return max_over_ranks(a.elapsed_time(b) / reps)

    # ---- BASELINE configs[2]: 4096 probes x N gallery, top-5.  One GPU: the whole gallery; N GPUs: sharded by identity,
    # probes split across the ranks, exchange + merge inside the timed region (dist.ShardedGallery).
    P3 = 4096
    # probes per SURVEY 8d: enrolled rows + N(0, sigma) noise (a true top-1 with a margin), every 7th a pure-random impostor
    g3 = torch.Generator(device=dev).manual_seed(77)          # same seed on every rank: the same probe set everywhere
    rows3 = torch.randint(0, N, (P3,), generator=g3, device=dev)
    probes3 = G[rows3] + 0.03 * torch.randn((P3, 512), generator=g3, device=dev) if world > 1 else None
    if probes3 is None:       # one GPU: the fp32 generator copy of the gallery is gone, regenerate the chosen rows' blocks
        probes3 = torch.empty((P3, 512), dtype=torch.float32, device=dev)
        gg = torch.Generator(device=dev).manual_seed(1234)
        for s0 in range(0, N, 1 << 18):
            blk = torch.randn((min(1 << 18, N - s0), 512), generator=gg, device=dev)
            blk = blk / blk.norm(dim=1, keepdim=True)
            sel = (rows3 >= s0) & (rows3 < s0 + blk.shape[0])
            probes3[sel] = blk[rows3[sel] - s0]
        probes3 += 0.03 * torch.randn((P3, 512), generator=g3, device=dev)
    probes3[::7] = torch.randn((probes3[::7].shape[0], 512), generator=g3, device=dev)
    sc3 = torch.empty((P3, topk), dtype=torch.float32, device=dev)
    ix3 = torch.empty((P3, topk), dtype=torch.int64, device=dev)
    ac3 = torch.empty((P3,), dtype=torch.uint8, device=dev)

    def match_all():
        ctx.frb_match(probes3.data_ptr(), P3, topk, thr, 1, sc3.data_ptr(), ix3.data_ptr(), ac3.data_ptr(), None, st)

    c3_1gpu_ms = timed(match_all, 10)
    ms3 = (C.c_float * 3)()
    c3_parts = []
    for _ in range(3):
        ctx.frb_match_profile(probes3.data_ptr(), P3, topk, thr, 1, sc3.data_ptr(), ix3.data_ptr(), ac3.data_ptr(), st, ms3)
        c3_parts.append(list(ms3))
    c3_parts = np.median(np.array(c3_parts), axis=0)
    p256_parts = []
    for i in range(4):
        ctx.frb_match_profile(emb.data_ptr(), B, topk, thr, 1, sc.data_ptr(), ix.data_ptr(), ac.data_ptr(), st, ms3)
        p256_parts.append(list(ms3))
    p256_parts = np.median(np.array(p256_parts[1:]), axis=0)
    sharded = None
    if world > 1:
        from facerecognitionpipeline_b200.dist import ShardedGallery, shard_bounds, split_probes
        match_all()
        torch.cuda.synchronize(dev)
        ix_ref = ix3.clone()
        ac_ref = ac3.clone()
        lo, hi = shard_bounds(N, world, rank)
        plo, phi = split_probes(P3, world, rank)
        mine = probes3[plo:phi].contiguous()
        res = {}
        sg_peer = None
        for mode in ("peer", "nccl"):
            try:
                sg = ShardedGallery(ctx=ctx, exchange=mode, max_probes=P3, max_k=8)
            except RuntimeError as e:          # no peer access between the GPUs: the NCCL exchange still runs
                res[mode] = dict(error=str(e)[:200])
                continue
            if mode == "peer":
                sg_peer = sg
            sg.upload_shard(G[lo:hi], N)
            o_sc, o_ix, o_ac = sg.match(mine, k=topk, thr=thr, n_probes=P3)
            torch.cuda.synchronize(dev)
            same = bool(torch.equal(o_ix, ix_ref) and torch.equal(o_ac, ac_ref))
            flag = torch.tensor([int(same)], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            ms = timed(lambda: sg.match(mine, k=topk, thr=thr, n_probes=P3), 20)
            res[mode] = dict(ms=ms, ids_identical_to_unsharded=bool(flag.item()),
                             rows_through_exact_fixup=int(ctx._lib.frb_match_last_flagged(ctx.handle)))
        # where the time goes (NCCL exchange, same steps timed one by one with events; max over ranks each)
        from facerecognitionpipeline_b200.dist import all_gather_balanced
        parts = None
        try:
            rec = torch.empty((P3, topk, 2), dtype=torch.int64, device=dev)
            every = torch.empty((world, P3, topk, 2), dtype=torch.int64, device=dev)
            evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(10)]
            barrier()
            for it in range(10):
                evs[it][0].record(stream)
                allp = all_gather_balanced(mine, P3)
                evs[it][1].record(stream)
                ctx.frb_match_packed(allp.data_ptr(), P3, topk, thr, 1, rec.data_ptr(), st)
                evs[it][2].record(stream)
                dist.all_gather_into_tensor(every.view(-1), rec.view(-1))
                ctx.frb_topk_merge_packed(every.data_ptr(), world, P3, topk, thr, sc3.data_ptr(), ix3.data_ptr(), ac3.data_ptr(), st)
                evs[it][3].record(stream)
            torch.cuda.synchronize(dev)
            barrier()
            tt = [float(np.mean([evs[it][j].elapsed_time(evs[it][j + 1]) for it in range(2, 10)])) for j in range(3)]
            parts = dict(probe_all_gather=max_over_ranks(tt[0]), local_match=max_over_ranks(tt[1]),
                         record_all_gather_and_merge=max_over_ranks(tt[2]),
                         note="event-separated steps of the NCCL exchange; a collective's time includes waiting for the slowest rank")
        except Exception as ex:
            parts = dict(error=str(ex)[:200])
        best = min((m for m in res if "ms" in res[m]), key=lambda m: res[m]["ms"])
        sharded = dict(workload=f"{P3} probes x {N} x 512 gallery, top-{topk}, gallery sharded by identity over {world} GPUs, "
                                f"probes split over the ranks; probe exchange + per-shard match + top-k exchange + merge timed",
                       exchange=best, sharded_match_ms=res[best]["ms"], sharded_probes_per_s=P3 / (res[best]["ms"] / 1e3),
                       one_gpu_match_ms=c3_1gpu_ms, sharded_vs_1gpu=c3_1gpu_ms / res[best]["ms"],
                       strong_scaling_efficiency=c3_1gpu_ms / res[best]["ms"] / world,
                       tflops=P3 * 1024.0 * N / (res[best]["ms"] / 1e3) / 1e12, by_exchange=res, nccl_parts_ms=parts)
        del G
        # back to the replicated gallery is not needed: everything below works on the embed path only

    # ---- BASELINE configs[3]: landmark warp -> IR-101 -> match at batch 1024 per GPU (data-parallel).  The match runs
    # against the gallery resident now (whole gallery at 1 GPU; this rank's shard after the sharded section).
    c4 = None
    if not args.no_c4:
        import cv2
        from facerecognitionpipeline_b200.face_recognition import estimate_matrix, similarity_template
        B4 = 1024
        r4 = np.random.default_rng(7 + rank)
        tpl = similarity_template(112)
        frames = np.stack([cv2.GaussianBlur(r4.integers(0, 256, (256, 256, 3), dtype=np.uint8), (0, 0), 2.0) for _ in range(16)])
        jobs = (_native.WarpJob * B4)()
        src_bytes = 0.0
        for i in range(B4):
            ang, scl = np.deg2rad(r4.uniform(-20, 20)), r4.uniform(1.5, 2.2)
            R = np.array([[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]]) * scl
            lm = ((tpl - 56) @ R.T + np.array([128 + r4.uniform(-8, 8), 128 + r4.uniform(-8, 8)]) + r4.normal(0, 0.5, (5, 2))).astype(np.float32)
            M = estimate_matrix(lm, tpl)
            jobs[i].src_off, jobs[i].H, jobs[i].W, jobs[i].pitch = (i % 16) * 256 * 256 * 3, 256, 256, 256 * 3
            for j, v in enumerate(np.asarray(M, np.float64).reshape(6)):
                jobs[i].M[j] = float(v)
            src_bytes += (112 * scl) ** 2 * 3
        d_frames = torch.from_numpy(frames).to(dev)
        x4 = torch.empty((B4, 112, 112, 3), dtype=torch.bfloat16, device=dev)
        emb4 = torch.empty((B4, 512), dtype=torch.float32, device=dev)
        sc4 = torch.empty((B4, topk), dtype=torch.float32, device=dev)
        ix4 = torch.empty((B4, topk), dtype=torch.int64, device=dev)
        ac4 = torch.empty((B4,), dtype=torch.uint8, device=dev)

        def c4_step():
            ctx.frb_warp_normalize(d_frames.data_ptr(), jobs, B4, 112, None, x4.data_ptr(), st)
            ctx.frb_embed(x4.data_ptr(), B4, flags, emb4.data_ptr(), None, None, st)
            ctx.frb_match(emb4.data_ptr(), B4, topk, thr, 1, sc4.data_ptr(), ix4.data_ptr(), ac4.data_ptr(), None, st)

        c4_ms = timed(c4_step, 5)
        warp_ms = timed(lambda: ctx.frb_warp_normalize(d_frames.data_ptr(), jobs, B4, 112, None, x4.data_ptr(), st), 10)
        warp_bytes = src_bytes + B4 * 75264.0
        c4 = dict(workload=f"BASELINE configs[3]: 5-landmark warp+normalise -> IR-101 -> top-{topk} match, batch {B4} per GPU, "
                           f"probe-dp{world}; warp sources 256x256 RGB frames resident in HBM, matrices from cv2.estimateAffinePartial2D on the host (untimed)",
                  faces_per_s=world * B4 / (c4_ms / 1e3), ms_per_step=c4_ms, batch_per_gpu=B4,
                  backbone_frac_of_burst=(flops_face * B4 / (c4_ms / 1e3) / 1e12) / peaks["tf_burst"],
                  warp=dict(ms=warp_ms, algorithmic_bytes=warp_bytes, gbs=warp_bytes / (warp_ms / 1e3) / 1e9,
                            frac_of_hbm=warp_bytes / (warp_ms / 1e3) / 1e9 / peaks["hbm"]))
        del x4, emb4

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel: per-layer CUDA-event durations of a few extra (untimed) steps through
    # frb_embed_profile; the kernel with the largest share of the embed section is reported (ALGORITHMIC flops of
    # its launches / their summed duration).  Event-separated launches do not overlap, so this is per-launch time.
    KNAMES = {0: "stem_tc_kernel", 1: "gemm2_sm100_kernel<64>", 2: "gemm2_sm100_kernel<128>", 3: "gemm2_sm100_kernel<256>",
              4: "conv_slab_sm100_kernel<64,1>", 5: "conv_slab_sm100_kernel<128,1>", 6: "conv_slab_sm100_kernel<128,2>",
              7: "gemm_sm100_kernel<256,0,1> + fc_finalize_kernel",
              8: "gemm2_multi_sm100_kernel<256> (persistent run of consecutive conv layers, one launch)",
              9: "conv_slab_multi_sm100_kernel (persistent run of identical slab conv layers, one launch)"}
    MAXL = 256
    ms = (C.c_float * MAXL)(); kid = (C.c_int * MAXL)(); fl = (C.c_double * MAXL)(); nl = C.c_int(0)
    per_kernel = {}
    PROF_STEPS = 5
    for i in range(PROF_STEPS + 1):
        ctx.frb_preprocess_u8(dev_crops[i % NBUF].data_ptr(), B, 112, x_bf16.data_ptr(), 0, st)
        ctx.frb_embed_profile(x_bf16.data_ptr(), B, flags, emb.data_ptr(), st, ms, kid, fl, MAXL, C.byref(nl))
        if i == 0:
            continue  # first profiled step creates the events
        for j in range(nl.value):
            acc = per_kernel.setdefault(abs(kid[j]), [0, 0.0, 0.0])
            acc[0] += 1 if kid[j] >= 0 else 0        # a negative id continues the launch reported before it
            acc[1] += ms[j]; acc[2] += fl[j]
    sched = (C.c_int * 4)()
    ctx.frb_embed_schedule(sched)
    schedule = dict(faces_per_pass=sched[0], front_layers=sched[1], front_images_per_sub_batch=sched[2], front_sub_batches=sched[3],
                    note="front = stem + the 112/56-pixel Cout-64 layers, launched sub-batch by sub-batch so their activations stay "
                         "in L2 between layers; the per_kernel times of those layers are sums over the sub-batches")
    dom = max(per_kernel, key=lambda k: per_kernel[k][1])
    n_l, ms_sum, fl_sum = per_kernel[dom]
    total_prof_ms = sum(v[1] for v in per_kernel.values())
    dom_tf = fl_sum / (ms_sum / 1e3) / 1e12
    section_tf = flops_face * B / (embed_ms / 1e3) / 1e12
    # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of that kernel from the committed `ncu --set full`
    # capture (profiles/r02_summary.md for the run, r01b_summary.md for the two single-layer kernels); null when the dominant kernel is not the one that was captured
    traffic = {3: 26.94e6, 6: 57.30e6, 8: RUN_TRAFFIC}.get(dom)
    roofline = dict(bound="tensor", achieved=dom_tf, peak=peaks["tf_sustained"], unit="TFLOP/s",
                    frac=dom_tf / peaks["tf_sustained"], traffic=traffic, kernel=KNAMES[dom],
                    launches_per_step=n_l // PROF_STEPS, avg_launch_us=ms_sum / n_l * 1e3,
                    share_of_embed=ms_sum / total_prof_ms,
                    backbone_section=dict(achieved=section_tf, frac=section_tf / peaks["tf_sustained"], ms=embed_ms,
                                          frac_of_burst=section_tf / peaks["tf_burst"]),
                    note=f"dominant kernel {KNAMES[dom]}: algorithmic {fl_sum / n_l / 1e9:.1f} GFLOP per launch / "
                         f"{ms_sum / n_l * 1e3:.1f} us average launch (CUDA events between layers on the launch stream, "
                         f"{PROF_STEPS} steps after the timed region); peak = bf16_tflops_sustained ({peaks['src']}). "
                         f"backbone_section = whole IR-101 embed ({flops_face / 1e9:.3f} GFLOP/face x {B} / {embed_ms:.3f} ms "
                         f"inside the timed steps, programmatic dependent launch on); match section {match_ms:.3f} ms/step")
    # every kernel of the step against ITS roofline (burst and sustained tensor peak, or the measured copy bandwidth)
    def tensor_entry(name, gflop, us, n):
        tf = gflop / 1e3 / (us / 1e6)
        # burst / sustained: the cuBLAS figures of MEASURED_PEAKS.json (a kernel timed alone can exceed the burst one:
        # it is a measurement of another GEMM, not a limit); nominal: 2250 TFLOP/s dense bf16 at 1.965 GHz
        return dict(kernel=name, bound="tensor", launches_per_step=n, us_per_step=us, algorithmic_gflop=gflop, achieved=tf,
                    unit="TFLOP/s", frac_of_burst=tf / peaks["tf_burst"], frac_of_sustained=tf / peaks["tf_sustained"],
                    frac_of_nominal=tf / 2250.0)

    def hbm_entry(name, nbytes, us, n, note=None):
        gbs = nbytes / 1e9 / (us / 1e6)
        d = dict(kernel=name, bound="hbm", launches_per_step=n, us_per_step=us, algorithmic_bytes=nbytes, achieved=gbs,
                 unit="GB/s", frac_of_hbm=gbs / peaks["hbm"])
        if note:
            d["note"] = note
        return d

    per_kernel_list = []
    for kid_, (n_, ms_, fl_) in sorted(per_kernel.items()):
        us = ms_ / PROF_STEPS * 1e3
        if kid_ == 0:      # stem: reads the bf16 crops, writes the 112x112x64 activation
            per_kernel_list.append(hbm_entry(KNAMES[0], B * (112 * 112 * 3 * 2 + 112 * 112 * 64 * 2.0), us, n_ // PROF_STEPS,
                                             f"also {fl_ / PROF_STEPS / 1e9:.1f} GFLOP on the tensor cores (K = 27 padded to 32)"))
        else:
            per_kernel_list.append(tensor_entry(KNAMES[kid_], fl_ / PROF_STEPS / 1e9, us, n_ // PROF_STEPS))
    # P = 256 sits on the ridge: one pass over the bf16 gallery (1.02 GB) and 256 x N x 1024 FLOP (0.27 PFLOP) each take
    # ~150-160 us at the measured peaks, so the entry carries both fractions
    e256 = hbm_entry("match_filter2_kernel (P = %d)" % B, N * 1024.0, float(p256_parts[1]) * 1e3, 1,
                     "ridge point: the bf16 gallery is streamed once (HBM) while the tensor cores do P x N x 1024 FLOP; "
                     "both roofline times are within 10 % of each other at P = 256")
    t256 = tensor_entry("", B * 1024.0 * N / 1e9, float(p256_parts[1]) * 1e3, 1)
    e256.update(algorithmic_gflop=t256["algorithmic_gflop"], tflops=t256["achieved"], frac_of_tensor_burst=t256["frac_of_burst"],
                frac_of_tensor_sustained=t256["frac_of_sustained"])
    per_kernel_list.append(e256)
    per_kernel_list.append(dict(kernel="probe_prepare_kernel", us_per_step=float(p256_parts[0]) * 1e3, bound="latency"))
    per_kernel_list.append(dict(kernel="match_finalize_kernel + exact fix-up kernels (device-side row list)",
                                us_per_step=float(p256_parts[2]) * 1e3, bound="latency"))
    per_kernel_list.append(tensor_entry("match_filter2_kernel (P = 4096, BASELINE configs[2] on one GPU)",
                                        P3 * 1024.0 * N / 1e9, float(c3_parts[1]) * 1e3, 1))
    if c4:
        per_kernel_list.append(hbm_entry("warp_normalize_kernel (1024 faces)", c4["warp"]["algorithmic_bytes"],
                                         c4["warp"]["ms"] * 1e3, 1, "source footprint + 75 264 B written per face (SURVEY 8d)"))
    # read-only and write-only HBM streams measured here (libfrb200's own probe, 1 GiB, 16-byte accesses): the yardstick for
    # kernels that only read (match filter) or only write (stem) next to the read+write copy figure of MEASURED_PEAKS.json
    stream = None
    try:
        scratch = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
        t_ms = C.c_float(0)
        ctx.frb_debug_stream_bw(scratch.data_ptr(), scratch.numel(), 0, 5, C.byref(t_ms))
        rd = scratch.numel() / t_ms.value / 1e6
        ctx.frb_debug_stream_bw(scratch.data_ptr(), scratch.numel(), 1, 5, C.byref(t_ms))
        wr = scratch.numel() / t_ms.value / 1e6
        del scratch
        stream = dict(read_only_gbs=rd, write_only_gbs=wr, copy_gbs=peaks["hbm"],
                      note="frb_debug_stream_bw, 1 GiB, measured in this run after the timed region")
        for e in per_kernel_list:
            if e.get("bound") == "hbm" and e["kernel"].startswith("match_filter"):
                e["frac_of_read_only_stream"] = e["achieved"] / rd
            if e.get("bound") == "hbm" and e["kernel"].startswith("stem"):
                e["frac_of_write_only_stream"] = e["achieved"] / wr
    except Exception as ex:   # the probe is informative only
        stream = dict(error=str(ex)[:200])
    roofline["stream_peaks"] = stream
    roofline["frac_of_burst"] = dom_tf / peaks["tf_burst"]
    roofline["per_kernel"] = per_kernel_list
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        ref = CpuReference(N)
        te, tm = ref.sample(64, 8)
        threads = ref.threads
        t_asis = ref.sample_as_is(2)
        cpu = dict(value=1.0 / (te + tm), unit=UNIT, cores=threads, kind="port", as_is_match_s_per_probe=t_asis,
                   as_is_value=1.0 / (te + t_asis),
                   sample=f"64 faces torch-eager fp32 IR-101 (batch 32, {te * 1e3:.1f} ms/face) + 8 probes numpy "
                          f"dot+argsort vs {N} x 512 f32 with the matrix cached ({tm * 1e3:.1f} ms/probe); 'as_is' = the "
                          f"reference's search() verbatim, which re-stacks the {N} templates on every call ({t_asis:.2f} s/probe, 2 probes)")
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=K, warmup=W, ms_per_step=ms_total / K,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16", data="synthetic",
                config=dict(workload=f"AdaFace IR-101 batch-{B}/GPU embed + L2-norm + top-5 match vs {N}-identity "
                                     f"gallery (BASELINE configs[1] + the metric's 1M gallery)",
                            batch_per_gpu=B, gallery_rows=N, top_k=topk, threshold=thr,
                            parallelism=f"probe-dp{world}, gallery+weights replicated, no data-path collective",
                            l2=f"per-step working set (~{N * 1024 / 1e9:.1f} GB bf16 gallery + 0.13 GB weights + >1 GB "
                               f"activations) exceeds the 126 MB L2; input crops rotate over {NBUF} distinct batches"),
                clocks=clocks,
                e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h),
                gpu_launches=int(launches), roofline=roofline, embed_ms=embed_ms, match_ms=match_ms, embed_schedule=schedule,
                match_4096=dict(workload=f"BASELINE configs[2] on ONE GPU: {P3} probes x {N} x 512, top-{topk}", ms=c3_1gpu_ms,
                                probes_per_s=P3 / (c3_1gpu_ms / 1e3), tflops=P3 * 1024.0 * N / (c3_1gpu_ms / 1e3) / 1e12,
                                frac_of_burst=P3 * 1024.0 * N / (c3_1gpu_ms / 1e3) / 1e12 / peaks["tf_burst"],
                                parts_ms=dict(prepare=float(c3_parts[0]), filter=float(c3_parts[1]), finalize=float(c3_parts[2]))))
    if sharded:
        line["sharded"] = sharded
        line["sharded_match_ms"] = sharded["sharded_match_ms"]
        line["sharded_probes_per_s"] = sharded["sharded_probes_per_s"]
        line["sharded_vs_1gpu"] = sharded["sharded_vs_1gpu"]
    if c4:
        line["c4"] = c4
    if cpu:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--gallery", type=int, default=1_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c4", action="store_true", help="skip the warp -> embed -> match batch-1024 sub-measurement")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
