"""Device-resident dynamic store and maintenance: id -> row hash, removal kernels, partition
surgery (split / delete / add, partition_manager.cpp:393-554) and the cost-model policy (maintenance_policies.cpp:33-202)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _qb():
    import quake_b200 as qb
    return qb


def _build(n=20000, d=32, nlist=40, metric="l2", seed=3):
    qb = _qb()
    torch.manual_seed(seed)
    x = torch.randn(n, d)
    bp = qb.IndexBuildParams()
    bp.nlist, bp.metric = nlist, metric
    idx = qb.QuakeIndex()
    idx.build(x, torch.arange(n, dtype=torch.int64), bp)
    return idx, x


def _exhaustive_equals_bruteforce(idx, x_by_id, q, k=10):
    """Search every partition: the answer must be the brute-force top-k over the vectors currently in the index."""
    qb = _qb()
    sp = qb.SearchParams()
    sp.k, sp.nprobe = k, idx.nlist()
    res = idx.search(q, sp)
    ids = torch.tensor(sorted(x_by_id.keys()), dtype=torch.int64)
    xs = torch.stack([x_by_id[int(i)] for i in ids])
    gt = ids[torch.cdist(q.double(), xs.double()).topk(k, largest=False).indices]
    assert torch.equal(res.ids, gt)


def test_remove_semantics_duplicates_absent_and_order():
    """remove(): duplicates and absent ids are ignored (std::set, partition_manager.cpp:306-310); each list ends up as
    the reference's swap-with-last loop leaves it (dynamic_inverted_list.cpp:137-149) -- replayed here on the host."""
    idx, x = _build()
    st = idx.store
    before = {int(p): st.get_list(int(p))[1].cpu().tolist() for p in st.partition_ids()}
    g = torch.Generator().manual_seed(1)
    rem = torch.randperm(20000, generator=g)[:3000]
    rem_dups = torch.cat([rem, rem[:500], torch.tensor([10 ** 7, 10 ** 7 + 1])])  # duplicates + ids that do not exist
    info = idx.remove(rem_dups)
    assert idx.ntotal() == 17000 and info.modify_count == rem_dups.numel()
    rs = set(rem.tolist())
    for p, ids in before.items():
        lst = list(ids)
        i = 0
        while i < len(lst):  # the reference's loop
            if lst[i] in rs:
                lst[i] = lst[-1]
                lst.pop()
            else:
                i += 1
        assert st.get_list(p)[1].cpu().tolist() == lst, f"partition {p}"
    # the id -> row table followed the moved rows
    left = torch.tensor(sorted(set(range(20000)) - rs), dtype=torch.int64)
    got = idx.get(left[::37])
    assert torch.equal(got.cpu(), x[left[::37]])
    with pytest.raises(RuntimeError):
        idx.get(rem[:3])
    idx.remove(torch.tensor([5, 5, 5]) + 0 * rem[:3])  # possibly absent, certainly duplicated: must not corrupt counts
    assert idx.ntotal() in (16999, 17000)
    assert sorted(idx.get_ids().tolist()) == sorted(set(left.tolist()) - {5})


def test_add_remove_interleaved_keeps_index_consistent():
    idx, x = _build(n=12000, nlist=30)
    by_id = {i: x[i] for i in range(12000)}
    g = torch.Generator().manual_seed(2)
    nxt = 12000
    for step in range(4):
        xa = torch.randn(1500, 32, generator=g)
        ida = torch.arange(nxt, nxt + 1500, dtype=torch.int64)
        idx.add(xa, ida)
        for j in range(1500):
            by_id[nxt + j] = xa[j]
        nxt += 1500
        live = torch.tensor(sorted(by_id.keys()), dtype=torch.int64)
        rem = live[torch.randperm(live.numel(), generator=g)[:900]]
        idx.remove(rem)
        for r in rem.tolist():
            del by_id[r]
        assert idx.ntotal() == len(by_id)
    assert sorted(idx.get_ids().tolist()) == sorted(by_id.keys())
    _exhaustive_equals_bruteforce(idx, by_id, torch.randn(25, 32, generator=g))


def test_split_delete_add_partitions():
    """partition_manager.cpp:393-554: split two partitions (k-means with K = 2 each), delete the originals, add the
    halves; then delete a partition with re-assignment. ntotal, membership and exhaustive search stay intact."""
    qb = _qb()
    idx, x = _build(n=15000, nlist=25)
    by_id = {i: x[i] for i in range(15000)}
    pids = torch.tensor([3, 11], dtype=torch.int64)
    sizes = [idx.store.size_of(3), idx.store.size_of(11)]
    cents, vecs, ids = idx.split_partitions(pids)
    assert cents.shape[0] == 4 and [int(v.shape[0]) for v in vecs][0] + int(vecs[1].shape[0]) == sizes[0]
    assert int(vecs[2].shape[0]) + int(vecs[3].shape[0]) == sizes[1]
    idx.delete_partitions(pids, reassign=False)
    assert idx.nlist() == 23 and idx.ntotal() == 15000 - sum(sizes)
    new = idx.add_partitions((cents, vecs, ids))
    assert new.tolist() == [25, 26, 27, 28] and idx.nlist() == 27 and idx.ntotal() == 15000
    assert idx.parent.ntotal() == 27
    # every half holds the vectors nearer to its own centroid than to its sibling's (k-means fixed point up to ties)
    for j in range(2):
        va, vb = vecs[2 * j][:, :32].cpu(), vecs[2 * j + 1][:, :32].cpu()
        ca, cb = cents[2 * j].cpu(), cents[2 * j + 1].cpu()
        assert float(((va - ca).norm(dim=1) <= (va - cb).norm(dim=1) + 1e-4).float().mean()) > 0.99
        assert float(((vb - cb).norm(dim=1) <= (vb - ca).norm(dim=1) + 1e-4).float().mean()) > 0.99
    idx.delete_partitions(torch.tensor([7], dtype=torch.int64), reassign=True)
    assert idx.nlist() == 26 and idx.ntotal() == 15000
    assert sorted(idx.get_ids().tolist()) == list(range(15000))
    _exhaustive_equals_bruteforce(idx, by_id, torch.randn(20, 32))
    # a fixed-nprobe search still works through the id -> slot table after the surgery
    sp = qb.SearchParams()
    sp.k, sp.nprobe = 5, 4
    r = idx.search(x[:50], sp)
    assert torch.equal(r.ids[:, 0], torch.arange(50))


class _ScriptedEstimator:
    """Stands in for MaintenanceCostEstimator (pinned against the reference in tests/test_maintenance_model.py): the
    decisions are scripted by partition size so that the test checks the policy's WIRING -- which partitions get
    deleted / split for given deltas, in which order, and what the index looks like afterwards."""

    def __init__(self):
        self.calls = []

    def compute_delete_delta(self, size, hit_rate, total, avg_rate, avg_size):
        self.calls.append(("delete", size, round(hit_rate, 3)))
        return -100.0 if size < 100 else 100.0

    def compute_split_delta(self, size, hit_rate, total):
        self.calls.append(("split", size, round(hit_rate, 3)))
        return -100.0 if (size > 2000 and hit_rate > 0.5) else 100.0

    def compute_delete_delta_w_reassign(self, size, hit_rate, total, counts, sizes, rates):
        self.calls.append(("reassign", size, sum(counts)))
        return -100.0


@pytest.mark.parametrize("rejection", [False, True])
def test_maintenance_policy_wiring(rejection):
    """maintenance_policies.cpp:33-172: window gate, hit rates from the recorded searches, delete-before-split order,
    deleted partitions' vectors re-assigned, split = k-means(2) + delete + add, local refinement, ntotal preserved."""
    qb = _qb()
    torch.manual_seed(4)
    d = 16
    centers = torch.randn(12, d) * 8
    sizes = [400] * 10 + [3000, 40]  # one oversized cluster, one tiny
    xs = torch.cat([centers[j] + torch.randn(sizes[j], d) for j in range(12)])
    n = xs.shape[0]
    # partitions = the generating clusters: sizes are known exactly
    offs = np.concatenate([[0], np.cumsum(sizes)])
    vecs = [xs[offs[j]:offs[j + 1]] for j in range(12)]
    ids = [torch.arange(offs[j], offs[j + 1], dtype=torch.int64) for j in range(12)]
    cents = torch.stack([v.mean(0) for v in vecs])
    idx = qb.QuakeIndex.from_partitions(cents, vecs, ids, "l2")
    assert idx.nlist() == 12 and idx.ntotal() == n
    pid_sizes = {int(p): idx.store.size_of(int(p)) for p in idx.store.partition_ids()}
    big, small = 10, 11
    mp = qb.MaintenancePolicyParams()
    mp.window_size, mp.refinement_radius, mp.refinement_iterations, mp.min_partition_size = 200, 3, 1, 32
    mp.enable_delete_rejection = rejection
    idx.initialize_maintenance_policy(mp)
    est = _ScriptedEstimator()
    idx.maintenance_policy._estimator = est
    sp = qb.SearchParams()
    sp.k, sp.nprobe = 5, 1
    info = idx.maintenance()
    assert info.n_splits == 0 and info.n_deletes == 0 and not est.calls  # window not full yet
    big_vecs, _ = idx.store.get_list(big)
    idx.search(big_vecs[:200].cpu() + 0.01, sp)  # 200 queries, all of them hit the big partition
    info = idx.maintenance()
    assert info.n_deletes == 1 and info.n_splits == 1
    assert any(c[0] == "reassign" for c in est.calls) == rejection  # the re-assignment-aware check runs only on demand
    hot = [c for c in est.calls if c[0] == "split" and c[1] == pid_sizes[big]]
    assert hot and hot[0][2] == 1.0  # hit rate of the big partition: 200 hits / window 200
    pids = idx.store.partition_ids().tolist()
    assert big not in pids and small not in pids and idx.nlist() == 12 - 1 - 1 + 2
    assert idx.parent.ntotal() == idx.nlist()
    assert idx.ntotal() == n and sorted(idx.get_ids().tolist()) == list(range(n))
    r = idx.search(xs[:64], sp)  # nearest-partition search still finds (almost) every vector itself
    assert float((r.ids[:, 0] == torch.arange(64)).float().mean()) > 0.9
    by_id = {i: xs[i] for i in range(n)}
    _exhaustive_equals_bruteforce(idx, by_id, torch.randn(10, d))


def test_gpu_latency_model_is_measured_and_monotone():
    from quake_b200 import maintenance as mt
    m = mt.ListScanLatencyEstimator(32, n_values=[1, 64, 4096, 65536], k_values=[1, 16], device=torch.device("cuda", 0))
    t = m.model
    assert t.shape == (4, 2) and np.all(t > 0)
    assert t[3, 0] > t[0, 0]  # scanning 65536 vectors costs more than scanning one
    assert m.estimate_scan_latency(70000, 10) > m.estimate_scan_latency(64, 10)
