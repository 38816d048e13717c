"""GPU parity of the QuakeIndex surface (build / search / add / remove / save / load / refit) against the
CPU oracle and the golden fixtures produced by the compiled reference."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as orc
from tests.conftest import GOLDEN

pytestmark = pytest.mark.gpu
REL_TOL = 1e-4


def _qb():
    import quake_b200 as qb
    return qb


def _assert_result(res_ids, res_dist, oi, od, exact_dist=True):
    ids, dist = res_ids.cpu().numpy(), res_dist.cpu().numpy()
    oi, od = np.asarray(oi), np.asarray(od)
    assert np.array_equal(ids, oi), f"ids differ in {(ids != oi).sum()} of {ids.size} slots"
    fin = np.isfinite(od)
    assert np.array_equal(dist[~fin], od[~fin])
    if exact_dist:
        assert np.array_equal(dist[fin], od[fin]), "distances not bit-identical"
    else:
        assert np.allclose(dist[fin], od[fin], rtol=REL_TOL, atol=0)


# ------------------------------------------------------------------ golden: the reference's own results
@pytest.mark.parametrize("metric", ["l2", "ip"])
@pytest.mark.parametrize("tag", ["serial_small", "serial", "batched", "k100", "all"])
def test_search_matches_reference_golden(metric, tag):
    """Load an index the REFERENCE built and saved, search it on the GPU, compare with the REFERENCE's own
    search results (tests/golden/search.npz)."""
    qb = _qb()
    S = np.load(os.path.join(GOLDEN, "search.npz"))
    idx = qb.QuakeIndex()
    idx.load(os.path.join(GOLDEN, f"index_{metric}"))
    nq, k, nprobe, batched = S[f"{metric}_{tag}_cfg"]
    sp = qb.SearchParams()
    sp.k, sp.nprobe, sp.batched_scan = int(k), int(nprobe), bool(batched)
    res = idx.search(torch.from_numpy(S[f"{metric}_q"][:nq]), sp)
    assert res.ids.dtype == torch.int64 and res.distances.dtype == torch.float32
    assert tuple(res.ids.shape) == (nq, k)
    _assert_result(res.ids, res.distances, S[f"{metric}_{tag}_ids"], S[f"{metric}_{tag}_dist"], exact_dist=not batched)


# ------------------------------------------------------------------ build + search vs oracle on the same index
@pytest.mark.parametrize("metric,d,nlist,k,nprobe", [("l2", 128, 64, 10, 8), ("ip", 128, 64, 10, 8),
                                                      ("l2", 96, 37, 10, 5), ("l2", 20, 16, 100, 16)])
def test_build_search_matches_oracle(metric, d, nlist, k, nprobe):
    qb = _qb()
    torch.manual_seed(1234)
    n = 20000
    x = torch.randn(n, d)
    ids = torch.arange(n, dtype=torch.int64)
    bp = qb.IndexBuildParams()
    bp.nlist, bp.metric = nlist, metric
    idx = qb.QuakeIndex()
    info = idx.build(x, ids, bp)
    assert idx.ntotal() == n and idx.nlist() == nlist and info.n_vectors == n
    assert sorted(idx.get_ids().tolist()) == list(range(n))
    torch.manual_seed(4321)
    for Q in (5, 64):  # < 20 and >= 20 queries: the reference's two coarse-scan arithmetic branches
        q = torch.randn(Q, d)
        sp = qb.SearchParams()
        sp.k, sp.nprobe = k, nprobe
        res = idx.search(q, sp)
        oi, od = orc.search_index_like(idx, q, k=k, nprobe=nprobe)
        _assert_result(res.ids, res.distances, oi.numpy(), od.numpy())


def test_flat_index_and_padding():
    """nlist <= 1: one partition, every query scans everything (quake_index.cpp:66-79); k > ntotal pads with
    -1 / +inf (query_coordinator.cpp:589-601); empty query batch gives empty tensors (:476-482)."""
    qb = _qb()
    torch.manual_seed(3)
    x = torch.randn(7, 16)
    idx = qb.QuakeIndex()
    idx.build(x, torch.arange(7) + 50, qb.IndexBuildParams())
    assert idx.nlist() == 1 and idx.parent is None
    sp = qb.SearchParams()
    sp.k = 10
    q = torch.randn(4, 16)
    res = idx.search(q, sp)
    oi, od = orc.search_index_like(idx, q, k=10, nprobe=1)
    _assert_result(res.ids, res.distances, oi.numpy(), od.numpy())
    assert (res.ids[:, 7:] == -1).all() and torch.isinf(res.distances[:, 7:]).all()
    empty = idx.search(torch.empty(0, 16), sp)
    assert empty.ids.numel() == 0 and empty.distances.numel() == 0
    # large flat index: recall vs brute force (test/cpp/search_recall_tests.cpp:160-189)
    x = torch.randn(30000, 32)
    idx = qb.QuakeIndex()
    idx.build(x, torch.arange(30000), qb.IndexBuildParams())
    q = torch.randn(50, 32)
    res = idx.search(q, sp)
    td, ti = torch.cdist(q, x).topk(10, largest=False)
    assert torch.equal(res.ids, ti)


def test_add_remove_search():
    """PartitionManager::add / remove semantics (partition_manager.cpp:123-320) + search afterwards."""
    qb = _qb()
    torch.manual_seed(11)
    d, n = 64, 8000
    x = torch.randn(n, d)
    idx = qb.QuakeIndex()
    bp = qb.IndexBuildParams()
    bp.nlist = 32
    idx.build(x, torch.arange(n), bp)
    xa = torch.randn(3000, d)
    ida = torch.arange(n, n + 3000)
    info = idx.add(xa, ida)
    assert info.modify_count == 3000 and idx.ntotal() == n + 3000
    # every added vector sits in the list of its nearest centroid
    pids, lists, cv, ci = orc.index_lists(idx)
    want = orc.assign(xa.numpy(), cv, "l2")
    member = {}
    for p, (v, i) in zip(pids, lists):
        for t in i:
            member[int(t)] = int(p)
    got = np.array([member[int(t)] for t in ida.tolist()])
    assert np.array_equal(got, ci[want])
    g = torch.Generator().manual_seed(99)
    rem = torch.randperm(n + 3000, generator=g)[:1500]
    idx.remove(rem)
    assert idx.ntotal() == n + 1500
    left = set(idx.get_ids().tolist())
    assert left == set(range(n + 3000)) - set(rem.tolist())
    got = idx.get(torch.tensor([n + 5, 3]) if (n + 5) in left and 3 in left else torch.tensor(sorted(left)[:2]))
    assert got.shape[1] == d
    q = torch.randn(40, d)
    sp = qb.SearchParams()
    sp.k, sp.nprobe = 10, 6
    res = idx.search(q, sp)
    oi, od = orc.search_index_like(idx, q, k=10, nprobe=6)
    _assert_result(res.ids, res.distances, oi.numpy(), od.numpy())
    with pytest.raises(RuntimeError):
        idx.add(torch.randn(2, d), torch.tensor([1 << 40, 5]))  # ids must be <= INT_MAX
    with pytest.raises(RuntimeError):
        idx.add(torch.randn(2, d), torch.tensor([7, 7]) + 100000)  # ids must be unique


def test_save_load_roundtrip(tmp_path):
    """Index directory in the reference's v3 format (dynamic_inverted_list.cpp:338-419): our writer ->
    the oracle's reader (same parser the golden tests use on reference-written files) and our reader."""
    qb = _qb()
    torch.manual_seed(5)
    x = torch.randn(5000, 24)
    idx = qb.QuakeIndex()
    bp = qb.IndexBuildParams()
    bp.nlist = 20
    idx.build(x, torch.arange(5000) * 2, bp)
    p = str(tmp_path / "idx")
    idx.save(p)
    m, pids, lists, cv, ci = orc.read_index_dir(p)
    a_pids, a_lists, a_cv, a_ci = orc.index_lists(idx)
    assert m == 1 and np.array_equal(pids, a_pids) and np.array_equal(cv, a_cv) and np.array_equal(ci, a_ci)
    for (v, i), (v2, i2) in zip(lists, a_lists):
        assert np.array_equal(v, v2) and np.array_equal(i, i2)
    idx2 = qb.QuakeIndex()
    idx2.load(p)
    q = torch.randn(30, 24)
    sp = qb.SearchParams()
    sp.k, sp.nprobe = 10, 4
    r1, r2 = idx.search(q, sp), idx2.search(q, sp)
    assert torch.equal(r1.ids, r2.ids) and torch.equal(r1.distances, r2.distances)


# ------------------------------------------------------------------ k-means pieces through the C ABI
@pytest.mark.parametrize("metric", ["l2", "ip"])
def test_kmeans_assign_update_match_oracle(metric):
    """qk_kmeans_assign == argmin/argmax in the reference's per-pair arithmetic, ties to the lowest index;
    qk_partition_by_assignment + qk_kmeans_accumulate == compute_centroids' data-order sums (bit-exact)."""
    from quake_b200 import clustering, _lib
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(21)
    n, K, d = 30000, 300, 48
    x = torch.randn(n, d, generator=g)
    c = torch.randn(K, d, generator=g)
    c[7] = c[3]  # exact tie between two centroids -> lowest index must win
    m = _lib.QK_METRIC_INNER_PRODUCT if metric == "ip" else _lib.QK_METRIC_L2
    xd, cd = clustering.pad_rows(x, dev), clustering.pad_rows(c, dev)
    a = clustering.assign_points(xd, d, cd, m)
    want = orc.assign(x.numpy(), c.numpy(), metric)
    assert np.array_equal(a.cpu().numpy(), want)
    assert not (a == 7).any()
    counts, offsets, order = clustering.partition_by_assignment(a, K)
    sums = clustering.centroid_sums(xd, d, order, offsets, K)
    wsum, wcnt = orc.centroid_sums(x.numpy(), want, K)
    assert np.array_equal(counts.cpu().numpy(), wcnt)
    assert np.array_equal(sums[:, :d].cpu().numpy(), wsum)
    o = order.cpu().numpy()
    offs = offsets.cpu().numpy()
    for cidx in (0, 3, 100, K - 1):
        seg = o[offs[cidx]:offs[cidx + 1]]
        assert np.array_equal(seg, np.nonzero(want == cidx)[0])  # ascending point index inside a list


@pytest.mark.parametrize("metric", ["l2", "ip"])
def test_kmeans_assign_filter_precision_policy(metric):
    """qk_kmeans_assign_filtered: the 2xTF32 filter returns the assignment of the 3xTF32 one (the winner is decided in
    exact arithmetic either way); many copies of one centroid make the proof fail for every point nearest to it --
    those points take the exact re-scan (lowest index still wins) and the run's policy falls back to three terms."""
    from quake_b200 import clustering, _lib
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(33)
    n, K, d = 20000, 256, 64
    x = torch.randn(n, d, generator=g)
    c = torch.randn(K, d, generator=g)
    m = _lib.QK_METRIC_INNER_PRODUCT if metric == "ip" else _lib.QK_METRIC_L2
    xd, cd = clustering.pad_rows(x, dev), clustering.pad_rows(c, dev)
    filt = clustering.AssignFilter(dev)
    assert filt.terms == 2
    a2 = clustering.assign_points(xd, d, cd, m, filt=filt)
    a3 = clustering.assign_points(xd, d, cd, m)
    want = orc.assign(x.numpy(), c.numpy(), metric)
    assert np.array_equal(a2.cpu().numpy(), want) and np.array_equal(a3.cpu().numpy(), want)
    filt.review()
    assert filt.terms == 2  # nothing (or next to nothing) was re-scanned on this data
    # seventeen copies each of centroids 5, 6 and 7: more tied rows than the refine keeps candidates, for > 0.5 % of the points
    c[16:32], c[32:48], c[48:64] = c[5], c[6], c[7]
    cd = clustering.pad_rows(c, dev)
    a2 = clustering.assign_points(xd, d, cd, m, filt=filt)
    want = orc.assign(x.numpy(), c.numpy(), metric)
    assert np.array_equal(a2.cpu().numpy(), want)
    tied = int(((a2 >= 5) & (a2 <= 7)).sum())
    assert not ((a2 >= 16) & (a2 < 64)).any() and tied > 0.005 * n
    rescanned = int(filt.stats[0].item())
    assert rescanned >= tied
    filt.review()
    assert filt.terms == 3


@pytest.mark.parametrize("metric", ["l2", "ip"])
@pytest.mark.parametrize("iters", [0, 3])
def test_kmeans_refine_matches_reference_golden(metric, iters):
    """kmeans_refine_partitions (clustering.cpp:99-182) vs the reference's own output."""
    from quake_b200 import clustering, _lib
    U = np.load(os.path.join(GOLDEN, "unit.npz"))
    dev = torch.device("cuda", 0)
    cents, vecs = U[f"rf_{metric}_cents"], U[f"rf_{metric}_vecs"]
    d = cents.shape[1]
    m = _lib.QK_METRIC_INNER_PRODUCT if metric == "ip" else _lib.QK_METRIC_L2
    cd = clustering.pad_rows(torch.from_numpy(cents), dev)
    vd = clustering.pad_rows(torch.from_numpy(vecs), dev)
    idd = torch.arange(vecs.shape[0], dtype=torch.int64, device=dev)
    c2, counts, nv, ni = clustering.kmeans_refine(cd, d, vd, idd, m, iters)
    assert np.array_equal(counts.cpu().numpy(), U[f"rf_{metric}_{iters}_sizes"])
    assert np.array_equal(ni.cpu().numpy(), U[f"rf_{metric}_{iters}_ids"])
    ref = U[f"rf_{metric}_{iters}_cents"]
    got = c2[:, :d].cpu().numpy()
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    ok = ~np.isnan(ref)
    assert np.allclose(got[ok], ref[ok], rtol=1e-5, atol=1e-6)


def test_refine_partitions_keeps_argmin_clustering():
    """test/cpp/partition_manager.cpp:121-167: refining a clustering that already is argmin(cdist) leaves
    every partition size unchanged; ntotal / nlist invariant after 3 refinement iterations."""
    qb = _qb()
    torch.manual_seed(8)
    x = torch.randn(6000, 32)
    idx = qb.QuakeIndex()
    bp = qb.IndexBuildParams()
    bp.nlist = 24
    idx.build(x, torch.arange(6000), bp)
    sizes0 = {int(p): idx.store.size_of(int(p)) for p in idx.store.partition_ids()}
    idx.refine_partitions(None, 0)
    sizes1 = {int(p): idx.store.size_of(int(p)) for p in idx.store.partition_ids()}
    assert sizes0 == sizes1
    idx.refine_partitions(None, 3)
    assert idx.ntotal() == 6000 and idx.nlist() == 24
    q = torch.randn(20, 32)
    sp = qb.SearchParams()
    sp.k, sp.nprobe = 10, 24
    res = idx.search(q, sp)
    td, ti = torch.cdist(q, x).topk(10, largest=False)
    assert torch.equal(res.ids, ti)


def test_merge_topk_matches_sort():
    """qk_merge_topk: S partial top-k lists -> one, ordered by (distance, id); -1 entries ignored."""
    import ctypes as C
    from quake_b200 import _lib
    from quake_b200._lib import ptr, check
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(4)
    S, Q, k = 8, 50, 10
    for metric in (_lib.QK_METRIC_L2, _lib.QK_METRIC_INNER_PRODUCT):
        d = torch.randn(S, Q, k, generator=g).abs()
        d[d < 0.3] = 0.25  # ties
        ids = torch.randperm(S * Q * k, generator=g).reshape(S, Q, k).to(torch.int64)
        ids[torch.rand(S, Q, k, generator=g) < 0.2] = -1
        ids[:, 7, :] = -1  # a query with no result at all
        dd, idd = d.to(dev).contiguous(), ids.to(dev).contiguous()
        oi = torch.empty(Q, k, dtype=torch.int64, device=dev)
        od = torch.empty(Q, k, dtype=torch.float32, device=dev)
        check(lib.qk_merge_topk(ptr(dd), ptr(idd), S, Q, k, metric, ptr(oi), ptr(od),
                                C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        torch.cuda.synchronize()
        desc = metric == _lib.QK_METRIC_INNER_PRODUCT
        for q in range(Q):
            pairs = [(float(d[s, q, j]), int(ids[s, q, j])) for s in range(S) for j in range(k) if ids[s, q, j] >= 0]
            pairs.sort(key=lambda t: ((-t[0]) if desc else t[0], t[1]))
            pairs = pairs[:k]
            want_i = [p[1] for p in pairs] + [-1] * (k - len(pairs))
            assert oi[q].tolist() == want_i
            got_d = od[q].cpu().numpy()
            assert np.array_equal(got_d[:len(pairs)], np.array([p[0] for p in pairs], np.float32))
            assert np.all(np.isinf(got_d[len(pairs):]))


# ------------------------------------------------------------------ APS (recall_target > 0)
@pytest.mark.parametrize("metric", ["l2", "ip"])
def test_aps_matches_reference_golden(metric):
    """Adaptive partition scanning through serial_scan (query_coordinator.cpp:521-579): the index the REFERENCE
    built, the REFERENCE's own APS results (recall_target 0.9, 10 % initial candidates, exact beta function)."""
    qb = _qb()
    S = np.load(os.path.join(GOLDEN, "search.npz"))
    idx = qb.QuakeIndex()
    idx.load(os.path.join(GOLDEN, f"index_{metric}"))
    sp = qb.SearchParams()
    sp.k, sp.recall_target, sp.initial_search_fraction, sp.use_precomputed = 10, 0.9, 0.1, False
    res = idx.search(torch.from_numpy(S[f"{metric}_q"]), sp)
    _assert_result(res.ids, res.distances, S[f"{metric}_aps_ids"], S[f"{metric}_aps_dist"])


@pytest.mark.parametrize("metric,pre,target,k", [("l2", True, 0.9, 10), ("l2", False, 0.8, 10), ("ip", True, 0.9, 10),
                                                 ("l2", True, 0.95, 100)])
def test_aps_matches_oracle(metric, pre, target, k):
    """Same index on both sides; ids, distances AND the number of partitions every query scanned must be those
    of the sequential loop."""
    qb = _qb()
    torch.manual_seed(1234)
    n, d, nlist = 30000, 32, 200
    x = torch.randn(n, d)
    if metric == "ip":
        x = x / x.norm(dim=1, keepdim=True)
    bp = qb.IndexBuildParams()
    bp.nlist, bp.metric = nlist, metric
    idx = qb.QuakeIndex()
    idx.build(x, torch.arange(n, dtype=torch.int64), bp)
    torch.manual_seed(4321)
    q = torch.randn(70, d)
    if metric == "ip":
        q = q / q.norm(dim=1, keepdim=True)
    sp = qb.SearchParams()
    sp.k, sp.recall_target, sp.initial_search_fraction, sp.use_precomputed = k, target, 0.25, pre
    res = idx.search(q, sp)
    pids, lists, cv, ci = orc.index_lists(idx)
    oi, od, cid, sc = orc.search_lists(pids, lists, cv, ci, q, k, 0, idx.metric, recall_target=target,
                                       initial_search_fraction=0.25, use_precomputed=pre, return_probe=True)
    _assert_result(res.ids, res.distances, oi.numpy(), od.numpy())
    assert np.array_equal(idx.last_partitions_scanned.cpu().numpy(), sc)
    assert 1 < sc.mean() < 50  # the early exit is actually exercised


def test_aps_needs_two_candidates():
    """compute_recall_profile throws with fewer than 2 candidate partitions (geometry.h:350-352)."""
    qb = _qb()
    torch.manual_seed(5)
    x = torch.randn(2000, 16)
    bp = qb.IndexBuildParams()
    bp.nlist = 8
    idx = qb.QuakeIndex()
    idx.build(x, torch.arange(2000, dtype=torch.int64), bp)
    sp = qb.SearchParams()
    sp.k, sp.recall_target, sp.initial_search_fraction = 5, 0.9, 0.01  # int(8 * 0.01) = 0 -> 1 candidate
    with pytest.raises(RuntimeError):
        idx.search(torch.randn(3, 16), sp)


# ------------------------------------------------------------------ sharded lists, simulated on one GPU
@pytest.mark.parametrize("metric", ["l2", "ip"])
def test_sharded_partials_merge_to_unsharded_result(metric):
    """the W per-shard partial top-k lists (lists owned round-robin by partition id), merged by
    qk_merge_topk, are bit-identical to the unsharded search. All W shards are held by one process here; the
    collective that moves the partials is covered by tests/test_sharded_gloo.py and test_gpu_sharded.py."""
    qb = _qb()
    from quake_b200 import clustering, sharded
    torch.manual_seed(1234)
    n, d, nlist, W = 30000, 96, 48, 4
    x = torch.randn(n, d)
    bp = qb.IndexBuildParams()
    bp.nlist, bp.metric = nlist, metric
    full = qb.QuakeIndex()
    full.build(x, torch.arange(n, dtype=torch.int64) + 3, bp)
    torch.manual_seed(4321)
    q = torch.randn(100, d)
    sp = qb.SearchParams()
    sp.k, sp.nprobe = 10, 12
    want = full.search(q, sp)
    xq = clustering.pad_rows(q, full.store.device)
    parts = []
    for r in range(W):
        sh = sharded.ShardedQuakeIndex(rank=r, world=W)
        sh.shard_from(full)
        assert sh.local.store.nlist == len(range(r, nlist, W))
        parts.append(sh.search_partial(xq, sp))
    pi = torch.stack([p[0] for p in parts]).contiguous()
    pd = torch.stack([p[1] for p in parts]).contiguous()
    mi, md = sharded.merge_partials_device(pd, pi, 10, full.metric)
    assert torch.equal(mi.cpu(), want.ids) and torch.equal(md.cpu(), want.distances)


def test_graph_replayed_search_equals_eager_and_tracks_updates():
    """Fixed-nprobe searches replay a CUDA graph captured per (batch size, k, nprobe, index version): same answer
    as the eager launches, results handed to the caller are copies (not the plan's static buffers), and a
    mutation of the index invalidates the plan."""
    qb = _qb()
    from quake_b200 import index as qidx
    torch.manual_seed(5)
    n, d = 20000, 64
    x = torch.randn(n, d)
    bp = qb.IndexBuildParams()
    bp.nlist = 40
    idx = qb.QuakeIndex()
    idx.build(x, torch.arange(n, dtype=torch.int64), bp)
    q = torch.randn(128, d)
    sp = qb.SearchParams()
    sp.k, sp.nprobe = 10, 6
    qidx.GRAPHS_ENABLED = False
    try:
        eager = idx.search(q, sp)
    finally:
        qidx.GRAPHS_ENABLED = True
    r1 = idx.search(q, sp)
    r2 = idx.search(torch.randn(128, d), sp)  # same plan, other queries
    r3 = idx.search(q, sp)
    assert torch.equal(r1.ids, eager.ids) and torch.equal(r1.distances, eager.distances)
    assert torch.equal(r3.ids, r1.ids) and not torch.equal(r2.ids, r1.ids)
    # mutation: the new vectors (copies of the queries) must be found by the next search
    idx.add(q.clone(), torch.arange(n, n + 128, dtype=torch.int64))
    r4 = idx.search(q, sp)
    assert torch.equal(r4.ids[:, 0], torch.arange(n, n + 128, dtype=torch.int64))
    assert float(r4.distances[:, 0].abs().max()) == 0.0
