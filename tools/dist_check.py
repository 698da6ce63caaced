#!/usr/bin/env python
"""Identity-sharded gallery over real NCCL (run under torchrun on >= 2 GPUs):
   sharded match (all-gather probes -> per-shard frb_match with global ids -> all-gather top-k -> frb_topk_merge)
   must equal the unsharded match of the same probes against the whole gallery, bit for bit in ids and accept
   flags.  Also times BASELINE config 3 (4096 probes x 1M x 512, top-5) on the N ranks.
   torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from facerecognitionpipeline_b200 import _native
from facerecognitionpipeline_b200.dist import ShardedGallery, shard_bounds, split_probes

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
ctx = _native.Context(local)
N, P, k, thr = int(os.environ.get("FRB_N", 1_000_000)), 4096, 5, 0.35

g = torch.Generator(device=dev).manual_seed(7)           # same seed on every rank -> same gallery everywhere
G = torch.randn((N, 512), generator=g, device=dev)
G /= G.norm(dim=1, keepdim=True)
probes = G[torch.randint(0, N, (P,), generator=g, device=dev)] + 0.03 * torch.randn((P, 512), generator=g, device=dev)
probes[::7] = torch.randn((probes[::7].shape[0], 512), generator=g, device=dev)   # impostors
st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

# unsharded answer on every rank
ctx.frb_gallery_upload(G.data_ptr(), N, 0, 1)
sc0 = torch.empty((P, k), dtype=torch.float32, device=dev); ix0 = torch.empty((P, k), dtype=torch.int64, device=dev)
ac0 = torch.empty((P,), dtype=torch.uint8, device=dev)
ctx.frb_match(probes.data_ptr(), P, k, thr, 1, sc0.data_ptr(), ix0.data_ptr(), ac0.data_ptr(), None, st)
torch.cuda.synchronize()

sg = ShardedGallery(ctx=ctx)
lo, hi = shard_bounds(N, world, rank)
sg.upload_shard(G[lo:hi].contiguous(), N)
plo, phi = split_probes(P, world, rank)
mine = probes[plo:phi].contiguous()
sc, ix, ac = sg.match(mine, k=k, thr=thr)
torch.cuda.synchronize()
ok = bool(torch.equal(ix, ix0) and torch.equal(ac, ac0) and (sc - sc0).abs().max().item() < 1e-6)

# timing: config 3
for _ in range(3):
    sg.match(mine, k=k, thr=thr)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
reps = 10
for _ in range(reps):
    sg.match(mine, k=k, thr=thr)
b.record(); torch.cuda.synchronize(); dist.barrier()
ms = torch.tensor([a.elapsed_time(b) / reps], device=dev, dtype=torch.float64)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
oks = torch.tensor([int(ok)], device=dev)
dist.all_reduce(oks, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps(dict(check="sharded==unsharded", ok=bool(oks.item()), world=world, N=N, P=P, k=k,
                          sharded_match_ms=float(ms.item()), probes_per_s=P / (float(ms.item()) / 1e3),
                          tflops=P * 1024.0 * N / (float(ms.item()) / 1e3) / 1e12)), flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
