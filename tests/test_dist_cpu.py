"""world_size-2 gloo tests (CPU) of the multi-GPU exchange logic in dist.py: shard bounds, the single padded
probe all-gather (uneven split, no count exchange), the single all-gather of packed (f64 score, i64 id) records and
the merge - with the oracle as local matcher/merger.  On GPUs the same exchange runs with frb_match_packed /
frb_topk_merge_packed over NCCL and as frb_match_sharded over peer memory: tests/test_gpu_dist.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from facerecognitionpipeline_b200 import dist as fd
from oracle import gallery as og


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 23, 100, 1_000_003):
        for w in (1, 2, 3, 8):
            spans = [fd.shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, N, P, k, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    G = rng.standard_normal((N, 512)).astype(np.float32)
    G /= np.linalg.norm(G, axis=1, keepdims=True)
    G[N // 2 + 3] = G[2]                      # exact cross-shard tie -> lower global id must win
    probes = rng.standard_normal((P, 512)).astype(np.float32)
    probes[0] = G[2]

    sg_holder = {}

    def local_match(pr, kk, thr, normalize):
        sg = sg_holder["sg"]
        idx, sc = og.search_batch(G[sg.lo:sg.hi], pr.numpy(), kk, normalize)
        idx = np.where(idx >= 0, idx + sg.lo, -1)
        return torch.from_numpy(sc), torch.from_numpy(idx)

    def merge(all_sc, all_ix, kk, thr):
        Gn, Pn = all_sc.shape[:2]
        sc = all_sc.numpy().transpose(1, 0, 2).reshape(Pn, -1)
        ix = all_ix.numpy().transpose(1, 0, 2).reshape(Pn, -1)
        o_s = np.full((Pn, kk), -np.inf)
        o_i = np.full((Pn, kk), -1, np.int64)
        for p in range(Pn):
            ok = ix[p] >= 0
            order = np.lexsort((ix[p][ok], -sc[p][ok]))[:kk]
            o_s[p, :len(order)] = sc[p][ok][order]
            o_i[p, :len(order)] = ix[p][ok][order]
        return torch.from_numpy(o_s.astype(np.float32)), torch.from_numpy(o_i), torch.from_numpy((o_s[:, 0] >= thr).astype(np.uint8))

    sg = fd.ShardedGallery(ctx=None, local_match=local_match, merge=merge)
    sg_holder["sg"] = sg
    lo, hi = fd.shard_bounds(N, world, rank)
    sg.upload_shard(G[lo:hi], N)
    plo, phi = fd.split_probes(P, world, rank)    # uneven split: exercises the var-len gather
    assert sg.exchange == "nccl"             # injected matcher -> the collective path (gloo here)
    sc, ix, ac = sg.match(torch.from_numpy(probes[plo:phi]), k=k, thr=0.5, n_probes=P)
    every = fd.all_gather_balanced(torch.from_numpy(probes[plo:phi]), P)
    assert torch.equal(every, torch.from_numpy(probes))
    eidx, esc = og.search_batch(G, probes, k)
    ok = bool(np.array_equal(ix.numpy(), eidx) and np.allclose(sc.numpy(), esc, atol=1e-6) and ix[0, 0].item() == 2
              and ix[0, 1].item() == N // 2 + 3 and ac[0].item() == 1)
    out[rank] = ok
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_sharded_gallery_exchange_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), 1001, 7, 5, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


def test_pack_records_roundtrip():
    sc = torch.tensor([[0.5, -float("inf")], [1.0 - 2 ** -52, 0.25]], dtype=torch.float64)
    ix = torch.tensor([[7, -1], [1 << 40, 3]], dtype=torch.int64)
    rec = fd.pack_records(sc, ix)
    assert rec.shape == (2, 2, 2) and rec.dtype == torch.int64 and rec.is_contiguous()
    s2, i2 = fd.unpack_records(rec)
    assert torch.equal(s2, sc) and torch.equal(i2, ix)
    # the record layout is the device's TopkRec: 8 bytes of f64 score, then 8 bytes of i64 id
    raw = rec.numpy().tobytes()
    assert np.frombuffer(raw[:8], np.float64)[0] == 0.5 and np.frombuffer(raw[8:16], np.int64)[0] == 7
