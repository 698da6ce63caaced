"""In-process multi-rank world for tests/test_gpu_dist.py (run as a SUBPROCESS, not collected by pytest).

Several ranks of an identity-sharded match live in this one process on one GPU: one ctx + one stream per rank, the
exchange buffers handed over as raw pointers (frb_xchg_connect_local).  The ranks wait for each other ON THE DEVICE,
so their kernels must really run concurrently and no enqueue may synchronise the context: the parent sets
CUDA_DEVICE_MAX_CONNECTIONS=32 (streams that alias onto one hardware queue would serialise a rank behind another rank's
wait kernel), CUDA_MODULE_LOADING=EAGER (a lazily loaded kernel's first launch synchronises the context; frb_xchg_create
also loads the kernels of the sharded path itself) and a short FRB_XCHG_TIMEOUT_MS; a
device-side timeout traps and kills only this subprocess.  Prints one JSON line: {"ok": bool, "cases": [...]}."""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import gallery as og  # noqa: E402


def _unit(x):
    return (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float32)


def problem(seed, N, P, dup=0):
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


def expect(G, probes, k, thr):
    eidx, esc = og.search_batch(G, probes, k)
    return eidx, esc, esc[:, 0].astype(np.float32) >= np.float32(thr)


class Rank:
    """One rank: own ctx, own stream, own exchange buffer."""

    def __init__(self, rank, world, max_probes, max_k):
        import torch
        from facerecognitionpipeline_b200 import _native
        self.rank, self.world = rank, world
        self.ctx = _native.Context(0)
        self.ctx.frb_xchg_create(world, rank, max_probes, max_k, (C.c_ubyte * 64)())   # also pre-allocates the workspaces
        self.stream = torch.cuda.Stream(torch.device("cuda", 0))


def make_world(world, max_probes=512, max_k=8):
    ranks = [Rank(r, world, max_probes, max_k) for r in range(world)]
    for a in ranks:
        for b in ranks:
            if a is not b:
                a.ctx.frb_xchg_connect_local(b.rank, b.ctx._lib.frb_xchg_local_buffer(b.ctx.handle))
    return ranks


def sharded_once(ranks, G, probes, k, thr):
    import torch
    from facerecognitionpipeline_b200.dist import shard_bounds, split_probes
    dev = torch.device("cuda", 0)
    N, P, world = len(G), len(probes), len(ranks)
    for r in ranks:
        lo, hi = shard_bounds(N, world, r.rank)
        if getattr(r, "shard", None) != (id(G), lo, hi):
            shard = np.ascontiguousarray(G[lo:hi])
            r.ctx.frb_gallery_upload(shard.ctypes.data if hi > lo else None, hi - lo, lo, 0)
            r.shard = (id(G), lo, hi)
    keep = []
    for r in ranks:
        plo, phi = split_probes(P, world, r.rank)
        mine = torch.from_numpy(np.ascontiguousarray(probes[plo:phi])).to(dev)
        keep.append((mine, torch.empty((P, k), dtype=torch.float32, device=dev), torch.empty((P, k), dtype=torch.int64, device=dev),
                     torch.empty((P,), dtype=torch.uint8, device=dev)))
    torch.cuda.synchronize()
    # Enqueue every rank's call back to back: nothing between two enqueues may synchronise the device (no allocation,
    # no copy: frb_xchg_create pre-allocated every workspace).
    for r, (mine, sc, ix, ac) in zip(ranks, keep):
        plo, phi = split_probes(P, world, r.rank)
        r.ctx.frb_match_sharded(mine.data_ptr(), plo, phi - plo, P, k, thr, 1, sc.data_ptr(), ix.data_ptr(), ac.data_ptr(),
                                C.c_void_p(r.stream.cuda_stream))
    torch.cuda.synchronize()
    return [(sc.cpu().numpy(), ix.cpu().numpy(), ac.cpu().numpy()) for _, sc, ix, ac in keep]


def main():
    from facerecognitionpipeline_b200 import _native
    cases = json.loads(sys.argv[1])
    report, all_ok = [], True
    worlds = {}
    for world, N, P, k, dup in cases:
        print(f"case world={world} N={N} P={P} k={k} dup={dup}", file=sys.stderr, flush=True)
        ranks = worlds.get(world) or worlds.setdefault(world, make_world(world))
        G, probes = problem(N + P, N, P, dup)
        thr, ok = 0.4, True
        eidx, esc, eacc = expect(G, probes, k, thr)
        for rep in range(3):                             # consecutive epochs: both slot parities, flags re-armed
            if rep == 2:
                probes = probes[::-1].copy()
                eidx, esc, eacc = expect(G, probes, k, thr)
            for ri, (sc, ix, ac) in enumerate(sharded_once(ranks, G, probes, k, thr)):
                fin = np.isfinite(esc)
                good = np.array_equal(ix, eidx) and bool(np.abs(sc[fin] - esc[fin]).max() <= 1e-6) \
                    and np.array_equal(ac.astype(bool), eacc)
                if not good and ok:        # first mismatch of the case: say what differs
                    bad = np.nonzero((ix != eidx).any(1))[0][:4]
                    print(f"  MISMATCH rep {rep} rank {ri}: rows {bad.tolist()} got {ix[bad].tolist()} want {eidx[bad].tolist()} "
                          f"scores {sc[bad].tolist()} want {esc[bad].tolist()} acc {ac[:8].tolist()} want {eacc[:8].astype(int).tolist()}",
                          file=sys.stderr, flush=True)
                ok = ok and good
        ok = ok and all(r.ctx.frb_xchg_status() == 0 for r in ranks)
        report.append(dict(case=[world, N, P, k, dup], ok=bool(ok)))
        all_ok = all_ok and ok
    # the same probes through frb_match on the whole gallery and through a 2-rank sharded match: identical bits
    ctx = _native.default_context(0)
    G, probes = problem(3, 40000, 300, 0)
    k, thr = 5, 0.35
    ctx.frb_gallery_upload(G.ctypes.data, len(G), 0, 0)
    sc0 = np.empty((300, k), np.float32); ix0 = np.empty((300, k), np.int64); ac0 = np.empty((300,), np.uint8)
    ctx.frb_match_host(probes.ctypes.data, 300, k, thr, 1, sc0.ctypes.data, ix0.ctypes.data, ac0.ctypes.data)
    ranks = worlds.get(2) or make_world(2)
    same = all(np.array_equal(ix, ix0) and np.array_equal(ac, ac0) and np.array_equal(sc, sc0)
               for sc, ix, ac in sharded_once(ranks, G, probes, k, thr))
    report.append(dict(case="sharded == unsharded frb_match (bitwise)", ok=bool(same)))
    print(json.dumps(dict(ok=bool(all_ok and same), cases=report)), flush=True)


if __name__ == "__main__":
    main()
