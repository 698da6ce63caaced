"""Multi-GPU partitioning of the embed-and-match path (one process per GPU, torch.distributed).

The reference is single-process / single-device; this module is new work with no reference
counterpart.  Two ways to partition, both from BASELINE.json's north star:

  * probe data-parallel: every rank holds the whole gallery and embeds+matches its own slice of the
    probe stream.  No data-path collective at all (weak scaling, `bench.py`'s default).
  * identity-sharded gallery: rank r holds gallery rows [lo_r, hi_r) and matches ALL probes against them
    with global row ids; the per-rank top-k lists are merged in canonical order (score desc, id asc).

Two exchanges implement the sharded match:

  * "peer" (default on GPUs): `frb_match_sharded` - no NCCL on the data path.  Every rank owns an exchange
    buffer mapped into all peers (cudaIpc over NVLink); the probe-prepare kernel stores the normalised
    probes into every rank's buffer, the finalize kernel stores each finished row's k records into every
    rank's result slot and the last row raises the flags, the merge kernel waits on the flags.  One
    enqueue per rank, no host synchronisation; torch.distributed is used once, to hand round the handles.
  * "nccl": ONE `all_gather_into_tensor` of the probes (deterministic balanced split, padded: no count
    exchange, no `.item()`), `frb_match_packed`, ONE `all_gather_into_tensor` of the 16-byte
    (f64 score, i64 id) records into a preallocated buffer, `frb_topk_merge_packed`.

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


def pack_records(sc64: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """(f64 score [P,k], i64 id [P,k]) -> [P,k,2] int64 records (score bits, id): the 16-byte layout of
    `TopkRec` on the device, so one collective moves both."""
    return torch.stack([sc64.contiguous().view(torch.int64), idx.contiguous()], dim=-1).contiguous()


def unpack_records(rec: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    return rec[..., 0].contiguous().view(torch.float64), rec[..., 1].contiguous()


def all_gather_balanced(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """Rows split with `split_probes` -> all n_total rows on every rank with ONE collective: every rank pads
    its slice to ceil(n_total / world) rows (known without communication), gathers into one buffer and
    drops the pad rows."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = split_probes(n_total, world, rank)
    if local.shape[0] != hi - lo:
        raise ValueError(f"rank {rank}: {local.shape[0]} local rows, split_probes({n_total}, {world}, {rank}) expects {hi - lo}")
    per = -(-int(n_total) // world)
    send = local.contiguous()
    if send.shape[0] != per:
        pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        pad[: send.shape[0]] = send
        send = pad
    out = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, send, group=group)
    if n_total == world * per:
        return out
    base, extra = divmod(int(n_total), world)   # ranks < extra hold `per` rows, the others per - 1
    keep = out.view(world, per, *local.shape[1:])
    return torch.cat([keep[:extra].reshape(-1, *local.shape[1:]), keep[extra:, :base].reshape(-1, *local.shape[1:])], dim=0)


class ShardedGallery:
    """Identity-sharded gallery over the ranks of a process group."""

    def __init__(self, ctx=None, group=None, local_match: Optional[Callable] = None, merge: Optional[Callable] = None,
                 exchange: str = "auto", max_probes: int = 4096, max_k: int = 32):
        self.ctx = ctx
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.lo = self.hi = 0
        self.n_total = 0
        self._local_match = local_match
        self._merge = merge
        self._bufs = {}
        if exchange not in ("auto", "peer", "nccl"):
            raise ValueError("exchange must be 'auto', 'peer' or 'nccl'")
        native = ctx is not None and local_match is None and merge is None
        self.exchange = "nccl"
        if native and exchange in ("auto", "peer"):
            ok = self._connect_peers(max_probes, max_k)
            if ok:
                self.exchange = "peer"
            elif exchange == "peer":
                raise RuntimeError("peer-memory exchange requested but the exchange buffers could not be mapped "
                                   f"on every rank: {self._peer_error}")
        self.max_probes, self.max_k = int(max_probes), int(max_k)

    # -- peer-memory exchange setup (cudaIpc handles handed round once over torch.distributed) -----------
    def _connect_peers(self, max_probes: int, max_k: int) -> bool:
        self._peer_error = ""
        dev = torch.device("cuda", self.ctx.device)
        handle = (C.c_ubyte * 64)()
        ok = 1
        if self.world > 8:
            ok, self._peer_error = 0, "more than 8 ranks"
        else:
            try:
                self.ctx.frb_xchg_create(self.world, self.rank, int(max_probes), int(max_k), handle)
            except Exception as e:   # noqa: BLE001 - reported through the collective vote below
                ok, self._peer_error = 0, str(e)
        mine = torch.tensor(list(bytes(handle)) + [ok], dtype=torch.uint8, device=dev)
        every = torch.empty((self.world, 65), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(every.view(-1), mine, group=self.group)
        every = every.cpu()
        if int(every[:, 64].min()) == 0:
            return False
        handles = (C.c_ubyte * (64 * self.world)).from_buffer_copy(every[:, :64].contiguous().numpy().tobytes())
        ok = 1
        try:
            self.ctx.frb_xchg_connect(handles)
        except Exception as e:   # noqa: BLE001
            ok, self._peer_error = 0, str(e)
        vote = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(vote, op=dist.ReduceOp.MIN, group=self.group)
        # every rank's buffer is mapped everywhere before anybody stores into a peer
        dist.barrier(group=self.group)
        return bool(vote.item())

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

    # -- buffers reused across calls (no allocation, no implicit synchronisation in the steady state) ----
    def _buf(self, name, shape, dtype, dev):
        t = self._bufs.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype or t.device != dev:
            t = torch.empty(shape, dtype=dtype, device=dev)
            self._bufs[name] = t
        return t

    def _outputs(self, P, k, dev):
        return (self._buf("sc32", (P, k), torch.float32, dev), self._buf("idx", (P, k), torch.int64, dev),
                self._buf("acc", (P,), torch.uint8, dev))

    # -- the exchange -----------------------------------------------------------------------
    def match(self, local_probes: torch.Tensor, k: int = 5, thr: float = 0.0, normalize: bool = True,
              probes_are_replicated: bool = False, n_probes: Optional[int] = None):
        """local_probes: this rank's rows `split_probes(P, world, rank)` of the P probes ([P_r, 512] fp32; e.g.
        the data-parallel embed output), or all P probes on every rank when probes_are_replicated.  n_probes
        = P (default: world * P_r, i.e. an even split).  Returns (scores f32 [P,k], ids i64 [P,k], accept u8 [P])
        for ALL P probes, identical on every rank.  The returned tensors are reused by the next call."""
        P = int(local_probes.shape[0]) if probes_are_replicated else int(n_probes if n_probes is not None else self.world * local_probes.shape[0])
        dev = local_probes.device
        if self.exchange == "peer":
            if P > self.max_probes or k > self.max_k:
                raise ValueError(f"{P} probes / k={k} exceed the exchange buffers ({self.max_probes} / {self.max_k})")
            lo, hi = split_probes(P, self.world, self.rank)
            mine = local_probes[lo:hi] if probes_are_replicated else local_probes
            if mine.shape[0] != hi - lo:
                raise ValueError(f"rank {self.rank}: {mine.shape[0]} local probes, split_probes expects {hi - lo}")
            mine = mine.contiguous()
            sc32, idx, acc = self._outputs(P, k, dev)
            st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            self.ctx.frb_match_sharded(mine.data_ptr(), lo, hi - lo, P, int(k), float(thr), 1 if normalize else 0,
                                       sc32.data_ptr(), idx.data_ptr(), acc.data_ptr(), st)
            return sc32, idx, acc
        probes = local_probes if probes_are_replicated else all_gather_balanced(local_probes, P, self.group)
        if self._local_match is not None:
            sc64, ix = self._local_match(probes, k, thr, normalize)          # [P,k] with global ids
            rec = pack_records(sc64, ix)
        else:
            rec = self._buf("rec", (P, k, 2), torch.int64, dev)
            st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            self.ctx.frb_match_packed(probes.contiguous().data_ptr(), P, int(k), float(thr), 1 if normalize else 0,
                                      rec.data_ptr(), st)
        every = self._buf("every", (self.world, P, k, 2), torch.int64, rec.device)
        dist.all_gather_into_tensor(every.view(-1), rec.view(-1), group=self.group)     # [G,P,k] records, ONE collective
        if self._merge is not None:
            all_sc, all_ix = unpack_records(every)
            return self._merge(all_sc, all_ix, k, thr)
        sc32, idx, acc = self._outputs(P, k, dev)
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        self.ctx.frb_topk_merge_packed(every.data_ptr(), self.world, P, int(k), float(thr), sc32.data_ptr(), idx.data_ptr(),
                                       acc.data_ptr(), st)
        return sc32, idx, acc
