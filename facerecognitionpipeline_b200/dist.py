"""Multi-GPU partitioning of the embed-and-match path (one process per GPU, torch.distributed).

The reference is single-process / single-device; this module is new work with no reference
counterpart.  Two ways to partition, both from BASELINE.json's north star:

  * probe data-parallel: every rank holds the whole gallery and embeds+matches its own slice of the
    probe stream.  No data-path collective at all (weak scaling, `bench.py`'s default).
  * identity-sharded gallery: rank r holds gallery rows [lo_r, hi_r).  Probes are all-gathered
    (P x 512 fp32), every rank matches ALL probes against its shard with global row ids
    (`frb_match` with first_global_id = lo_r), the per-rank top-k lists (f64 score, i64 id) are
    all-gathered over NCCL/NVLink (G x P x k x 16 bytes — latency-bound, not bandwidth-bound) and every
    rank merges them with the canonical order (`frb_topk_merge`).

`local_match` / `merge` are injectable so the world_size-2 gloo tests can drive the exchange logic on
CPU with the oracle as the checker; the defaults are the CUDA entry points.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n_rows: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced identity shards: the first n_rows % world ranks get one extra row."""
    base, extra = divmod(int(n_rows), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def split_probes(n_probes: int, world: int, rank: int) -> Tuple[int, int]:
    return shard_bounds(n_probes, world, rank)


def _all_gather_rows(t: torch.Tensor, group=None) -> torch.Tensor:
    """all_gather of equally shaped tensors -> stacked [world, ...]."""
    world = dist.get_world_size(group)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t.contiguous(), group=group)
    return torch.stack(out)


def _all_gather_varlen(t: torch.Tensor, group=None) -> torch.Tensor:
    """all_gather along dim 0 when ranks hold different row counts (pads to the max)."""
    world = dist.get_world_size(group)
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    m = max(counts)
    pad = torch.zeros((m,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)


class ShardedGallery:
    """Identity-sharded gallery over the ranks of a process group."""

    def __init__(self, ctx=None, group=None, local_match: Optional[Callable] = None, merge: Optional[Callable] = None):
        self.ctx = ctx
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.lo = self.hi = 0
        self.n_total = 0
        self._local_match = local_match or self._native_match
        self._merge = merge or self._native_merge

    # -- gallery residency ------------------------------------------------------------------
    def upload_shard(self, shard_rows, n_total: int):
        """shard_rows: this rank's rows [lo, hi) as a float32 array/tensor (host or device)."""
        self.n_total = int(n_total)
        self.lo, self.hi = shard_bounds(n_total, self.world, self.rank)
        assert len(shard_rows) == self.hi - self.lo, "shard size does not match shard_bounds"
        self._shard = shard_rows
        if self.ctx is not None:
            if isinstance(shard_rows, torch.Tensor) and shard_rows.is_cuda:
                self.ctx.frb_gallery_upload(shard_rows.data_ptr(), len(shard_rows), self.lo, 1)
            else:
                arr = np.ascontiguousarray(np.asarray(shard_rows), dtype=np.float32)
                self.ctx.frb_gallery_upload(arr.ctypes.data, len(arr), self.lo, 0)

    # -- default (CUDA) backends ------------------------------------------------------------
    def _native_match(self, probes: torch.Tensor, k: int, thr: float, normalize: bool):
        P = probes.shape[0]
        dev = probes.device
        sc32 = torch.empty((P, k), dtype=torch.float32, device=dev)
        idx = torch.empty((P, k), dtype=torch.int64, device=dev)
        acc = torch.empty((P,), dtype=torch.uint8, device=dev)
        sc64 = torch.empty((P, k), dtype=torch.float64, device=dev)
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        self.ctx.frb_match(probes.data_ptr(), P, k, float(thr), 1 if normalize else 0, sc32.data_ptr(), idx.data_ptr(),
                           acc.data_ptr(), sc64.data_ptr(), st)
        return sc64, idx

    def _native_merge(self, all_sc64: torch.Tensor, all_idx: torch.Tensor, k: int, thr: float):
        G, P = all_sc64.shape[0], all_sc64.shape[1]
        dev = all_sc64.device
        sc32 = torch.empty((P, k), dtype=torch.float32, device=dev)
        idx = torch.empty((P, k), dtype=torch.int64, device=dev)
        acc = torch.empty((P,), dtype=torch.uint8, device=dev)
        sc64 = torch.empty((P, k), dtype=torch.float64, device=dev)
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        self.ctx.frb_topk_merge(all_sc64.contiguous().data_ptr(), all_idx.contiguous().data_ptr(), G, P, k, float(thr),
                                sc32.data_ptr(), idx.data_ptr(), acc.data_ptr(), sc64.data_ptr(), st)
        return sc32, idx, acc

    # -- the exchange -----------------------------------------------------------------------
    def match(self, local_probes: torch.Tensor, k: int = 5, thr: float = 0.0, normalize: bool = True,
              probes_are_replicated: bool = False):
        """local_probes: this rank's [P_r, 512] fp32 probes (data-parallel embed output), or the full
        probe set on every rank when probes_are_replicated.  Returns (scores f32 [P,k], ids i64 [P,k],
        accept u8 [P]) for ALL P probes, identical on every rank; rank r's own probes are rows
        split_probes(P, world, r) when the probes were split with split_probes."""
        probes = local_probes if probes_are_replicated else _all_gather_varlen(local_probes, self.group)
        sc64, idx = self._local_match(probes, k, thr, normalize)          # [P,k] with global ids
        all_sc = _all_gather_rows(sc64, self.group)                        # [G,P,k]
        all_ix = _all_gather_rows(idx, self.group)
        return self._merge(all_sc, all_ix, k, thr)
