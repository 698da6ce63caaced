"""Identity-sharded gallery on real GPUs (new work, the reference is single-device).

  * one GPU: `frb_match_sharded` with a world of 1, and worlds of 2-3 ranks living in one (sub)process (one ctx + one
    stream per rank, exchange buffers handed over as raw pointers): probe push, flag waits, row push from the finalize
    and fix-up kernels, parity/epoch handling over repeated calls, merge - against the unsharded `frb_match` and the
    oracle, bit-exact ids;
  * >= 2 GPUs (skipped otherwise): one process per GPU, NCCL process group: the peer-memory exchange (cudaIpc) and the
    NCCL exchange (one packed all-gather) must both equal the unsharded match.
"""
import os
import socket

import numpy as np
import pytest

from oracle import gallery as og

pytestmark = pytest.mark.gpu


def _unit(x):
    return (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float32)


def _problem(seed, N, P, dup=0):
    rng = np.random.default_rng(seed)
    G = _unit(rng.standard_normal((N, 512)))
    t = min(11, N // 2 - 1)
    G[N - 1] = G[t]                                     # exact tie across shards: the lower global id must win
    if dup:
        where = np.sort(rng.choice(N, dup, replace=False))
        G[where] = G[where[0]]                          # defeats the filter proof -> exact fix-up rows
    probes = _unit(G[rng.integers(0, N, P)] + 0.04 * rng.standard_normal((P, 512)))
    probes[::5] = rng.standard_normal((len(probes[::5]), 512)) * 1.7     # impostors, unnormalised
    probes[1] = G[t]
    return G, probes


def _expect(G, probes, k, thr):
    eidx, esc = og.search_batch(G, probes, k)
    return eidx, esc, esc[:, 0].astype(np.float32) >= np.float32(thr)


CASES = [(1, 9000, 70, 5, 0), (2, 9001, 70, 5, 0), (3, 30000, 257, 3, 0), (2, 20000, 100, 5, 150), (2, 5000, 33, 4, 0),
         (3, 7, 10, 5, 0)]


@pytest.mark.timeout(600)
def test_sharded_match_in_process_world():
    """frb_match_sharded for worlds of 1-3 ranks on ONE GPU (tests/sharded_inproc_worker.py, a subprocess: the ranks
    wait for each other on the device, a timeout there traps).  (2, 5000): shards below 4096 rows take the dense exact
    path and the plain row push; (3, 7): a rank with 2 rows; dup = 150: rows fixed up by the exact kernels are pushed
    from there; last line of the report: bitwise equality with the unsharded frb_match."""
    import json
    import subprocess
    import sys
    env = dict(os.environ, CUDA_DEVICE_MAX_CONNECTIONS="32", CUDA_MODULE_LOADING="EAGER", FRB_XCHG_TIMEOUT_MS="8000")
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sharded_inproc_worker.py")
    proc = subprocess.run([sys.executable, worker, json.dumps(CASES)], env=env, capture_output=True, text=True, timeout=540)
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-4000:]
    rep = json.loads(proc.stdout.strip().splitlines()[-1])
    assert rep["ok"], rep


# ---------------------------------------------------------------------------------------------- real multi-GPU
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _nccl_worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    from facerecognitionpipeline_b200 import _native
    from facerecognitionpipeline_b200.dist import ShardedGallery, shard_bounds, split_probes
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    dev = torch.device("cuda", rank)
    ctx = _native.Context(rank)
    N, P, k, thr = 200_003, 1001, 5, 0.35
    G, probes = _problem(17, N, P, 0)
    eidx, esc, eacc = _expect(G[:1], probes[:1], 1, thr)      # oracle import check only; the full answer comes from frb_match
    ctx.frb_gallery_upload(G.ctypes.data, N, 0, 0)
    sc0 = np.empty((P, k), np.float32); ix0 = np.empty((P, k), np.int64); ac0 = np.empty((P,), np.uint8)
    ctx.frb_match_host(probes.ctypes.data, P, k, thr, 1, sc0.ctypes.data, ix0.ctypes.data, ac0.ctypes.data)
    lo, hi = shard_bounds(N, world, rank)
    plo, phi = split_probes(P, world, rank)
    mine = torch.from_numpy(np.ascontiguousarray(probes[plo:phi])).to(dev)
    ok = {}
    for mode in ("peer", "nccl"):
        sg = ShardedGallery(ctx=ctx, exchange=mode, max_probes=2048, max_k=8) if mode == "peer" else \
            ShardedGallery(ctx=ctx, exchange="nccl")
        sg.upload_shard(np.ascontiguousarray(G[lo:hi]), N)
        good = sg.exchange == mode
        for _ in range(3):
            sc, ix, ac = sg.match(mine, k=k, thr=thr, n_probes=P)
            torch.cuda.synchronize()
            good = good and np.array_equal(ix.cpu().numpy(), ix0) and np.array_equal(ac.cpu().numpy(), ac0) \
                and np.array_equal(sc.cpu().numpy(), sc0)
        ok[mode] = bool(good)
    out[rank] = ok
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_sharded_gallery_over_nccl_and_peer_memory():
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_nccl_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {r: {"peer": True, "nccl": True} for r in range(world)}
