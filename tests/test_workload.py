"""Workload harness (quake_b200/workload.py; reference: src/python/workload_generator.py): host logic on CPU, the
generate -> replay round trip on the GPU."""
import json

import numpy as np
import pytest
import torch

from quake_b200 import workload as wl


def test_compute_recall_and_knn_cpu():
    ids = torch.tensor([[1, 2, 3], [4, 5, 6]])
    gt = torch.tensor([[3, 2, 9], [7, 8, 9]])
    assert wl.compute_recall(ids, gt, 3).tolist() == pytest.approx([2 / 3, 0.0])
    g = torch.Generator().manual_seed(0)
    x = torch.randn(500, 8, generator=g)
    q = torch.randn(7, 8, generator=g)
    i, d = wl.knn(q, x, 5, "l2", chunk=128)
    ti = torch.cdist(q, x).topk(5, largest=False)
    assert torch.equal(i, ti.indices) and torch.allclose(d, ti.values, atol=1e-4)
    i2, d2 = wl.knn(q, x, 3, "ip", chunk=100)
    assert torch.equal(i2, (q @ x.T).topk(3).indices)


def test_samplers():
    torch.manual_seed(1)
    pool = torch.arange(100, 200)
    s = wl.UniformSampler().sample(pool, 10)
    assert s.numel() == 10 and len(set(s.tolist())) == 10 and all(100 <= v < 200 for v in s.tolist())
    cents = torch.tensor([[0.0, 0.0], [1.0, 0.0], [5.0, 0.0], [9.0, 0.0]])
    assign = torch.tensor([0] * 10 + [1] * 10 + [2] * 10 + [3] * 10)
    sm = wl.StratifiedClusterSampler(assign, cents)
    sm.update_ranks(0)
    got = sm.sample(torch.arange(40), 15, update_ranks=False)
    # nearest-first from cluster 0: all of cluster 0, then 5 of cluster 1
    assert set(range(10)) <= set(got.tolist()) and all(v < 20 for v in got.tolist()) and got.numel() == 15


@pytest.mark.gpu
def test_generate_and_replay_roundtrip(tmp_path):
    import quake_b200 as qb
    torch.manual_seed(5)
    base = torch.randn(6000, 16)
    queries = torch.randn(300, 16)
    gen = wl.DynamicWorkloadGenerator(tmp_path / "w", base, "l2", insert_ratio=0.3, delete_ratio=0.2, query_ratio=0.5,
                                      update_batch_size=200, query_batch_size=50, number_of_operations=14,
                                      initial_size=2500, cluster_size=250, cluster_sample_distribution="skewed",
                                      queries=queries, seed=11)
    book = gen.generate_workload()
    assert gen.workload_exists() and book["summary"]["n_operations"] >= 10
    on_disk = json.load(open(tmp_path / "w" / "runbook.json"))
    assert on_disk["parameters"]["initial_size"] == 2500 and len(on_disk["operations"]) == book["summary"]["n_operations"]
    mp = qb.MaintenancePolicyParams()
    mp.window_size = 50
    ev = wl.WorkloadEvaluator(tmp_path / "w", tmp_path / "out")
    res = ev.evaluate_workload("quake_b200", wl.QuakeWrapper(), {"nc": 10, "metric": "l2"}, {"k": 10, "nprobe": 10},
                               do_maintenance=True, m_params=mp, batch=True)
    assert len(res) == len(on_disk["operations"])
    # the resident set the index holds follows the runbook (maintenance never changes ntotal) ...
    for r in res:
        assert r["n_total"] == r["n_resident"]
    # ... and a search that probes every initial partition's worth of lists finds the exact neighbours
    recalls = [r["recall"] for r in res if r["operation_type"] == "query"]
    assert recalls and min(recalls) > 0.9
    assert (tmp_path / "out" / "quake_b200_results.json").exists()
    assert ev.summary["avg_query_recall"] == pytest.approx(float(np.mean(recalls)))


# ------------------------------------------------------------------ pinned against the reference's own generator
def _golden_workload():
    import os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    return np.load(os.path.join(here, "workload.npz")), json.load(open(os.path.join(here, "workload.json")))


def _golden_dataset(seed, n, nq, d):
    # the dataset of tests/golden/make_golden_workload.py
    g = torch.Generator().manual_seed(seed)
    centers = torch.randn(12, d, generator=g) * 4.0
    base = centers[torch.randint(0, 12, (n,), generator=g)] + torch.randn(n, d, generator=g)
    queries = centers[torch.randint(0, 12, (nq,), generator=g)] + torch.randn(nq, d, generator=g)
    return base, queries


class _RecordedClustering:
    """Stands in for the clustered index: the reference's own centroids (the generator only asks for those)."""

    def __init__(self, centroids):
        self._c = centroids

    def centroids(self):
        return self._c


@pytest.mark.parametrize("case", ["skewed", "uniform"])
def test_generator_draws_the_reference_stream(case, tmp_path):
    """The operation stream of DynamicWorkloadGenerator equals the one the REFERENCE's generator drew on the same
    inputs (tests/golden/workload.{npz,json}, written by make_golden_workload.py from
    /root/reference/src/python/workload_generator.py): operation types (np.random.choice), sampled ids of every
    operation and of the initial resident set (torch.randperm / randint inside the samplers, in the reference's call
    order), resident-set sizes, and the ground-truth neighbours of every query batch. The reference's clustering
    (assignments, centroids, and the state its k-means left torch's global generator in) is injected, so the test needs
    no GPU."""
    z, meta = _golden_workload()
    m = meta[case]
    base, queries = _golden_dataset(*m["dataset"])

    class Gen(wl.DynamicWorkloadGenerator):
        def initialize_clustered_index(self):
            self.assignments = torch.from_numpy(z[f"{case}_assignments"])
            # the reference's C++ k-means draws from torch's global generator: the state it left is part of the recording
            torch.set_rng_state(torch.from_numpy(z[f"{case}_rng_after_clustering"]))
            return _RecordedClustering(torch.from_numpy(z[f"{case}_centroids"]))

    gen = Gen(tmp_path / "w", base, queries=queries, **m["kwargs"])
    book = gen.generate_workload()
    assert book["summary"] == m["summary"]
    assert book["parameters"] == m["parameters"]
    assert torch.equal(torch.load(tmp_path / "w" / "initial_indices.pt"), torch.from_numpy(z[f"{case}_initial"]))
    ops = book["operations"]
    assert [ops[i]["type"] for i in sorted(ops)] == m["types"]
    assert [ops[i]["sample_size"] for i in sorted(ops)] == m["sizes"]
    assert [ops[i]["n_resident"] for i in sorted(ops)] == z[f"{case}_n_resident"].tolist()
    offs = z[f"{case}_op_offsets"]
    for j, i in enumerate(sorted(ops)):
        ids = torch.load(tmp_path / "w" / "operations" / f"{i}.pt")
        assert ids.tolist() == z[f"{case}_op_ids"][offs[j]:offs[j + 1]].tolist(), f"operation {i}"
        if ops[i]["type"] == "query":
            gt = torch.load(tmp_path / "w" / "operations" / f"{i}_gt_ids.pt")
            assert torch.equal(gt[:, :10], torch.from_numpy(z[f"{case}_gt_{i}"])), f"ground truth of operation {i}"
