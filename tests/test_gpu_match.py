"""Parity of the CUDA gallery match (frb_match: bf16 tcgen05 filter + exact f64 re-score, or exact scan for
small galleries) against oracle/gallery.py and the reference's own search outputs (golden).
Bar: top-k indices bit-exact after the canonical tie-break, accept/reject identical, |score| within 1e-6."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import gallery as og

pytestmark = pytest.mark.gpu

SCORE_TOL = 1e-6


def _unit(x):
    return (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float32)


def _match(ctx, G, probes, k, thr=0.35, normalize=1, first=0):
    G = np.ascontiguousarray(G, np.float32)
    probes = np.ascontiguousarray(probes, np.float32)
    P = len(probes)
    ctx.frb_gallery_upload(G.ctypes.data if len(G) else None, len(G), first, 0)
    sc = np.zeros((P, k), np.float32)
    ix = np.zeros((P, k), np.int64)
    ac = np.zeros((P,), np.uint8)
    ctx.frb_match_host(probes.ctypes.data, P, k, thr, normalize, sc.ctypes.data, ix.ctypes.data, ac.ctypes.data)
    return sc, ix, ac


def _check(G, probes, k, sc, ix, ac, thr, first=0):
    eidx, esc = og.search_batch(G, probes, k)
    eidx = np.where(eidx >= 0, eidx + first, -1)
    assert np.array_equal(ix, eidx)
    fin = np.isfinite(esc)
    # f32-rounded scores: absolute 1e-6 for |score| <= 1, relative beyond (unnormalised gallery rows)
    assert (np.abs(sc[fin] - esc[fin]) <= SCORE_TOL * np.maximum(1.0, np.abs(esc[fin]))).all()
    assert np.array_equal(ac.astype(bool), esc[:, 0].astype(np.float32) >= np.float32(thr))


def test_golden_reference_search_23_identities(ctx, golden_dir):
    """The reference's own GalleryManager.search outputs on the shipped adaface_ir_101 gallery."""
    z = np.load(os.path.join(golden_dir, "gallery_backups.npz"))
    s = np.load(os.path.join(golden_dir, "search_cases.npz"))
    tag = str(s["gallery_tag"])
    G, ids = z[tag + "/tpl"], z[tag + "/ids"]
    sc, ix, ac = _match(ctx, G, s["probes"], 5, thr=0.5)
    for p in range(len(s["probes"])):
        assert [str(ids[i]) for i in ix[p]] == [str(x) for x in s["ids"][p]]
        assert np.abs(sc[p] - s["scores"][p]).max() <= 2e-6
    _check(G, s["probes"], 5, sc, ix, ac, 0.5)


def test_verify_enrollment_on_all_shipped_galleries(ctx, golden_dir):
    z = np.load(os.path.join(golden_dir, "gallery_backups.npz"))
    for tag in sorted({k.split("/")[0] for k in z.files}):
        G = z[tag + "/tpl"]
        probes = z[tag + "/emb"].reshape(-1, 512)
        sc, ix, ac = _match(ctx, G, probes, 3, thr=0.5)
        assert np.array_equal(ix[:, 0], np.repeat(np.arange(23), 8))
        _check(G, probes, 3, sc, ix, ac, 0.5)


@pytest.mark.parametrize("N,P,k", [(1, 3, 5), (4, 9, 5), (100, 32, 5), (257, 130, 3), (4095, 64, 5),
                                   (4096, 64, 5), (4097, 5, 1), (70001, 300, 5), (300000, 129, 8)])
def test_synthetic_vs_oracle(ctx, N, P, k):
    rng = np.random.default_rng(N + P)
    G = _unit(rng.standard_normal((N, 512)))
    probes = _unit(G[rng.integers(0, N, P)] + 0.04 * rng.standard_normal((P, 512)))
    probes[P // 2:] = rng.standard_normal((P - P // 2, 512)) * 2.5      # impostors, unnormalised
    sc, ix, ac = _match(ctx, G, probes, k, thr=0.4)
    _check(G, probes, k, sc, ix, ac, 0.4)
    assert ac[: P // 2].all()


def test_unnormalised_gallery_rows_are_used_as_is(ctx):
    """search() normalises the query but NOT the gallery (gallery_manager.py:195-196)."""
    rng = np.random.default_rng(7)
    G = rng.standard_normal((9000, 512)).astype(np.float32) * rng.uniform(0.2, 3.0, (9000, 1)).astype(np.float32)
    probes = rng.standard_normal((40, 512)).astype(np.float32)
    sc, ix, ac = _match(ctx, G, probes, 5, thr=0.4)
    _check(G, probes, 5, sc, ix, ac, 0.4)


def test_ties_and_duplicates(ctx):
    rng = np.random.default_rng(9)
    for N in (64, 20000):
        G = _unit(rng.standard_normal((N, 512)))
        G[N - 5] = G[3]
        G[N // 2] = G[3]
        G[10] = G[11]
        probes = np.stack([G[3], G[10], G[7]])
        sc, ix, ac = _match(ctx, G, probes, 5, thr=0.9)
        assert ix[0, :3].tolist() == [3, N // 2, N - 5]
        assert ix[1, :2].tolist() == [10, 11]
        _check(G, probes, 5, sc, ix, ac, 0.9)


def test_threshold_is_greater_or_equal(ctx):
    """accept uses >= (face_matcher.py:205): a top-1 score exactly equal to the threshold accepts."""
    G = np.zeros((8, 512), np.float32)
    G[np.arange(8), np.arange(8)] = 1.0
    q = np.zeros((2, 512), np.float32)
    q[0, 2] = 1.0                      # score exactly 1.0 after normalisation (1/(1+1e-8) rounds to 1.0 in f32)
    q[1, 5] = 0.5
    q[1, 6] = 0.5
    sc, ix, ac = _match(ctx, G, q, 2, thr=float(np.float32(1.0 / (1.0 + 1e-8))))
    assert ix[0, 0] == 2 and ac[0] == 1 and ac[1] == 0


def test_empty_gallery_and_k_larger_than_n(ctx):
    q = np.ones((2, 512), np.float32)
    sc, ix, ac = _match(ctx, np.zeros((0, 512), np.float32), q, 3)
    assert (ix == -1).all() and not ac.any()
    G = _unit(np.random.default_rng(1).standard_normal((3, 512)))
    sc, ix, ac = _match(ctx, G, q, 5)
    assert (ix[:, 3:] == -1).all() and (ix[:, :3] >= 0).all() and np.isneginf(sc[:, 3:]).all()


def test_first_global_id_offsets_results(ctx):
    rng = np.random.default_rng(3)
    G = _unit(rng.standard_normal((6000, 512)))
    probes = _unit(rng.standard_normal((10, 512)))
    sc, ix, ac = _match(ctx, G, probes, 5, first=1_000_000)
    _check(G, probes, 5, sc, ix, ac, 0.35, first=1_000_000)


def test_sharded_merge_equals_unsharded(ctx):
    """Identity-sharded gallery on one GPU: match each shard with global ids, merge with frb_topk_merge
    (the kernel the NCCL path runs after its all-gather) == matching the whole gallery."""
    import torch
    rng = np.random.default_rng(21)
    N, P, k, shards = 50000, 70, 5, 3
    G = _unit(rng.standard_normal((N, 512)))
    G[40000] = G[100]                                   # tie across shards
    probes = _unit(G[rng.integers(0, N, P)] + 0.05 * rng.standard_normal((P, 512)))
    probes[0] = G[100]
    from facerecognitionpipeline_b200.dist import shard_bounds
    dev = torch.device("cuda", 0)
    pr = torch.from_numpy(probes).to(dev)
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    all_sc, all_ix = [], []
    for r in range(shards):
        lo, hi = shard_bounds(N, shards, r)
        ctx.frb_gallery_upload(np.ascontiguousarray(G[lo:hi]).ctypes.data, hi - lo, lo, 0)
        s32 = torch.empty((P, k), dtype=torch.float32, device=dev)
        s64 = torch.empty((P, k), dtype=torch.float64, device=dev)
        ix = torch.empty((P, k), dtype=torch.int64, device=dev)
        ac = torch.empty((P,), dtype=torch.uint8, device=dev)
        ctx.frb_match(pr.data_ptr(), P, k, 0.4, 1, s32.data_ptr(), ix.data_ptr(), ac.data_ptr(), s64.data_ptr(), st)
        all_sc.append(s64)
        all_ix.append(ix)
    A, I = torch.stack(all_sc).contiguous(), torch.stack(all_ix).contiguous()
    o32 = torch.empty((P, k), dtype=torch.float32, device=dev)
    oix = torch.empty((P, k), dtype=torch.int64, device=dev)
    oac = torch.empty((P,), dtype=torch.uint8, device=dev)
    ctx.frb_topk_merge(A.data_ptr(), I.data_ptr(), shards, P, k, 0.4, o32.data_ptr(), oix.data_ptr(), oac.data_ptr(), None, st)
    torch.cuda.synchronize()
    _check(G, probes, k, o32.cpu().numpy(), oix.cpu().numpy(), oac.cpu().numpy(), 0.4)
    assert oix[0, :2].tolist() == [100, 40000]


def test_full_size_1m_gallery_4096_probes_properties(ctx):
    """BASELINE config 3 shape on one GPU (4096 probes x 1M x 512, top-5): checked through properties that do
    not need a 4096 x 1M CPU matmul — planted rows come back at rank 1 with the planted score, rows are sorted,
    indices are valid and unique — plus an oracle check on a 96-probe subset."""
    import torch
    N, P, k = 1_000_000, 4096, 5
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(5)
    G = torch.randn((N, 512), generator=g, device=dev)
    G = G / G.norm(dim=1, keepdim=True)
    ctx.frb_gallery_upload(G.data_ptr(), N, 0, 1)
    rows = torch.randint(0, N, (P,), generator=g, device=dev)
    probes = G[rows] + 0.03 * torch.randn((P, 512), generator=g, device=dev)
    probes[P // 2:] = torch.randn((P - P // 2, 512), generator=g, device=dev)
    probes = probes / (probes.norm(dim=1, keepdim=True) + 1e-8)
    s32 = torch.empty((P, k), dtype=torch.float32, device=dev)
    ix = torch.empty((P, k), dtype=torch.int64, device=dev)
    ac = torch.empty((P,), dtype=torch.uint8, device=dev)
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    ctx.frb_match(probes.data_ptr(), P, k, 0.5, 0, s32.data_ptr(), ix.data_ptr(), ac.data_ptr(), None, st)
    torch.cuda.synchronize()
    assert ctx._lib.frb_match_last_flagged(ctx.handle) <= 8          # the filter's proof almost never fails
    h = P // 2
    assert torch.equal(ix[:h, 0], rows[:h])
    planted = (G[rows[:h]].double() * probes[:h].double()).sum(1)
    assert (s32[:h, 0].double() - planted).abs().max().item() <= SCORE_TOL
    assert ac[:h].all() and not ac[h:].any()
    assert (s32[:, :-1] >= s32[:, 1:]).all()
    assert (ix >= 0).all() and (ix < N).all()
    assert all(len(set(r)) == k for r in ix[::37].tolist())
    sub = torch.arange(0, P, P // 96, device=dev)[:96]
    Gh, ph = G.cpu().numpy(), probes[sub].cpu().numpy()
    eidx, esc = og.search_batch(Gh, ph, k, normalize=False)
    assert np.array_equal(ix[sub].cpu().numpy(), eidx)
    assert np.abs(s32[sub].cpu().numpy() - esc).max() <= SCORE_TOL


def test_config5_10m_identity_gallery_sharded_equals_unsharded(ctx):
    """BASELINE config 5 gallery size (10M identities x 512) on one GPU: the gallery is generated on the device, matched
    whole, then as two identity shards with global ids merged by frb_topk_merge (what the NCCL path does after its
    all-gather): ids and accept flags must be identical, planted rows come back at rank 1 with their exact f64 score."""
    import torch
    N, P, k, thr = 10_000_000, 256, 5, 0.5
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(17)
    G = torch.empty((N, 512), dtype=torch.float32, device=dev)
    for s0 in range(0, N, 1 << 20):
        blk = torch.randn((min(1 << 20, N - s0), 512), generator=g, device=dev)
        G[s0:s0 + blk.shape[0]] = blk / blk.norm(dim=1, keepdim=True)
    rows = torch.randint(0, N, (P,), generator=g, device=dev)
    probes = G[rows] + 0.03 * torch.randn((P, 512), generator=g, device=dev)
    probes[P // 2:] = torch.randn((P - P // 2, 512), generator=g, device=dev)
    probes = (probes / probes.norm(dim=1, keepdim=True)).contiguous()
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    def match(first):
        s32 = torch.empty((P, k), dtype=torch.float32, device=dev)
        s64 = torch.empty((P, k), dtype=torch.float64, device=dev)
        ix = torch.empty((P, k), dtype=torch.int64, device=dev)
        ac = torch.empty((P,), dtype=torch.uint8, device=dev)
        ctx.frb_match(probes.data_ptr(), P, k, thr, 0, s32.data_ptr(), ix.data_ptr(), ac.data_ptr(), s64.data_ptr(), st)
        torch.cuda.synchronize()
        return s32, s64, ix, ac

    ctx.frb_gallery_upload(G.data_ptr(), N, 0, 1)
    assert ctx._lib.frb_gallery_size(ctx.handle) == N
    w32, w64, wix, wac = match(0)
    h = P // 2
    assert torch.equal(wix[:h, 0], rows[:h])
    planted = (G[rows[:h]].double() * probes[:h].double()).sum(1)
    assert (w64[:h, 0] - planted).abs().max().item() <= 1e-12
    assert wac[:h].all() and not wac[h:].any()
    assert (w32[:, :-1] >= w32[:, 1:]).all() and (wix >= 0).all() and (wix < N).all()
    parts = []
    for lo, hi in ((0, N // 2), (N // 2, N)):
        ctx.frb_gallery_upload(G[lo:hi].data_ptr(), hi - lo, lo, 1)
        parts.append(match(lo))
    A = torch.stack([p[1] for p in parts]).contiguous()
    I = torch.stack([p[2] for p in parts]).contiguous()
    o32 = torch.empty((P, k), dtype=torch.float32, device=dev)
    oix = torch.empty((P, k), dtype=torch.int64, device=dev)
    oac = torch.empty((P,), dtype=torch.uint8, device=dev)
    ctx.frb_topk_merge(A.data_ptr(), I.data_ptr(), 2, P, k, thr, o32.data_ptr(), oix.data_ptr(), oac.data_ptr(), None, st)
    torch.cuda.synchronize()
    assert torch.equal(oix, wix) and torch.equal(oac, wac) and torch.equal(o32, w32)
    ctx.frb_gallery_upload(G[:256].data_ptr(), 256, 0, 1)      # release the 10M-row copies held by the context
    del G
    torch.cuda.empty_cache()


@pytest.mark.parametrize("n_hit", [1, 5, 16, 17, 120])
def test_flagged_rows_are_fixed_on_the_device(ctx, n_hit):
    """Rows whose filter proof fails are re-done by the exact fix-up kernels, which take the row list and its length
    from the device (no D2H + synchronise inside frb_match).  200 identical copies of one row defeat the proof (the
    re-scored survivors tie with everything the filter kept below them).  Up to 16 flagged rows are each scanned by a
    multiple of the 74 gallery partitions (one flagged row: the whole grid); above that the fix-up walks its list in
    strides (16 rows in flight in the partial kernel, 64 in the merge kernel)."""
    rng = np.random.default_rng(77)
    N, P, k = 30000, 150, 5
    G = _unit(rng.standard_normal((N, 512)))
    dup = np.sort(rng.choice(N, 200, replace=False))
    G[dup] = G[dup[0]]
    probes = _unit(rng.standard_normal((P, 512)))
    hit = np.sort(rng.choice(P, n_hit, replace=False))
    probes[hit] = _unit(G[dup[0]][None] + 0.02 * rng.standard_normal((len(hit), 512)))
    sc, ix, ac = _match(ctx, G, probes, k, thr=0.4)
    flagged = ctx._lib.frb_match_last_flagged(ctx.handle)
    assert flagged >= len(hit), flagged
    assert (ix[hit] == dup[:k][None]).all()            # exact ties: the lowest ids win
    _check(G, probes, k, sc, ix, ac, 0.4)


@pytest.mark.parametrize("N,k", [(500, 50), (20000, 33), (20000, 100)])
def test_top_k_above_32(ctx, N, k):
    """search() accepts any top_k (gallery_manager.py:197); k > 32 takes the dense exact scan."""
    rng = np.random.default_rng(N + k)
    G = _unit(rng.standard_normal((N, 512)))
    probes = _unit(rng.standard_normal((7, 512)))
    probes[0] = G[5]
    sc, ix, ac = _match(ctx, G, probes, k, thr=0.4)
    _check(G, probes, k, sc, ix, ac, 0.4)


def test_gallery_manager_search_any_top_k(ctx, tmp_path):
    from facerecognitionpipeline_b200.gallery_manager import GalleryManager
    rng = np.random.default_rng(5)
    gm = GalleryManager(gallery_path=str(tmp_path / "g.pkl"))
    E = _unit(rng.standard_normal((40, 512)))
    for i in range(40):
        gm.add_student(f"STU{i:04d}", f"n{i}", E[i:i + 1])
    res = gm.search(E[7], top_k=100)                    # more than the gallery holds: min(top_k, N) rows, as numpy slicing
    assert len(res) == 40 and res[0][0] == "STU0007"
    s = np.dot(E, E[7] / (np.linalg.norm(E[7]) + 1e-8))
    order = np.lexsort((np.arange(40), -s))
    assert [r[0] for r in res] == [f"STU{i:04d}" for i in order]
    assert len(gm.search(E[7], top_k=35)) == 35


def test_match_enqueues_without_synchronising_and_is_graph_capturable(ctx):
    """frb_match only enqueues (the per-call D2H + cudaStreamSynchronize is gone): once its workspaces
    exist the whole call can be captured into a CUDA graph and replayed on new probe contents."""
    import torch
    rng = np.random.default_rng(31)
    N, P, k = 20000, 96, 5
    G = _unit(rng.standard_normal((N, 512)))
    dev = torch.device("cuda", 0)
    ctx.frb_gallery_upload(G.ctypes.data, N, 0, 0)
    pr = torch.empty((P, 512), dtype=torch.float32, device=dev)
    s32 = torch.empty((P, k), dtype=torch.float32, device=dev)
    ix = torch.empty((P, k), dtype=torch.int64, device=dev)
    ac = torch.empty((P,), dtype=torch.uint8, device=dev)
    side = torch.cuda.Stream(dev)

    def enqueue(stream):
        ctx.frb_match(pr.data_ptr(), P, k, 0.4, 1, s32.data_ptr(), ix.data_ptr(), ac.data_ptr(), None,
                      C.c_void_p(stream.cuda_stream))

    probes0 = _unit(rng.standard_normal((P, 512)))
    pr.copy_(torch.from_numpy(probes0))
    with torch.cuda.stream(side):
        enqueue(side)                                   # warm-up: allocates the workspaces
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        enqueue(side)
    for seed in (1, 2):
        probes = _unit(G[np.random.default_rng(seed).integers(0, N, P)] + 0.05 * np.random.default_rng(seed).standard_normal((P, 512)))
        pr.copy_(torch.from_numpy(probes))
        torch.cuda.synchronize()
        graph.replay()
        torch.cuda.synchronize()
        _check(G, probes, k, s32.cpu().numpy(), ix.cpu().numpy(), ac.cpu().numpy(), 0.4)


@pytest.mark.parametrize("direction", ["down", "up"])
def test_exact_under_coherent_bf16_rounding(ctx, direction):
    """The filter's error bound is ||q - bf16(q)|| max||g|| + ||bf16(q)|| max||g - bf16(g)|| (+ accumulation): measured
    rounding distances instead of the element-wise worst case.  Worst case FOR that bound: every element sits just inside
    a bf16 rounding boundary on the same side, so all 512 rounding errors of a row have the sign of the element and the
    error of a score adds up coherently instead of averaging out.  Clusters of near-duplicate gallery rows 1e-4 apart
    in score make the approximate ranking wrong inside the cluster; the answer must still be the exact one."""
    rng = np.random.default_rng(11 if direction == "down" else 12)
    N, P, k = 24000, 96, 5

    def edge(x):   # move every element to the bf16 value nearest below |x|, then 0.49 ulp away from it on one side
        b = np.ascontiguousarray(x, np.float32).view(np.uint32) & np.uint32(0xFFFF0000)
        lo = b.view(np.float32)
        ulp = np.abs((b + np.uint32(0x00010000)).view(np.float32) - lo)
        return (lo + np.sign(lo) * ulp * (0.49 if direction == "down" else 0.51)).astype(np.float32)

    G = _unit(rng.standard_normal((N, 512)))
    base = rng.choice(N, 40, replace=False)
    for b in base:                                  # 30 near-duplicates of each base row
        dup = rng.choice(N, 30, replace=False)
        G[dup] = G[b] + 2e-4 * rng.standard_normal((30, 512)).astype(np.float32)
    G = edge(G)                                     # not re-normalised: the rows keep their boundary values
    probes = edge(_unit(G[base[rng.integers(0, 40, P)]] + 0.05 * rng.standard_normal((P, 512)).astype(np.float32)))
    sc, ix, ac = _match(ctx, G, probes, k, thr=0.4, normalize=0)       # probes used as they are: boundary values on both sides
    eidx, esc = og.search_batch(G, probes, k, normalize=False)
    assert np.array_equal(ix, eidx)
    assert (np.abs(sc - esc) <= SCORE_TOL * np.maximum(1.0, np.abs(esc))).all()
    assert np.array_equal(ac.astype(bool), esc[:, 0].astype(np.float32) >= np.float32(0.4))
    assert ctx._lib.frb_match_last_flagged(ctx.handle) < P            # the proof still passes for most rows
    # and through the normalising path (the device rounds bf16(q / ||q||): generic probe values, boundary gallery)
    sc, ix, ac = _match(ctx, G, probes, k, thr=0.4, normalize=1)
    _check(G, probes, k, sc, ix, ac, 0.4)
