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
# dram__bytes_read.sum + dram__bytes_write.sum of the persistent 66-layer run at batch 256 (ncu --set full of one launch, profiles/r01d_summary.md)
RUN_TRAFFIC = 305.98e6 + 1145.59e6
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
    del G

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

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = load_peaks()
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
    dom = max(per_kernel, key=lambda k: per_kernel[k][1])
    n_l, ms_sum, fl_sum = per_kernel[dom]
    total_prof_ms = sum(v[1] for v in per_kernel.values())
    dom_tf = fl_sum / (ms_sum / 1e3) / 1e12
    section_tf = flops_face * B / (embed_ms / 1e3) / 1e12
    # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of that kernel from the committed `ncu --set full`
    # capture (profiles/r01b_summary.md); null when the dominant kernel is not the one that was captured
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
                gpu_launches=int(launches), roofline=roofline, embed_ms=embed_ms, match_ms=match_ms)
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
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
