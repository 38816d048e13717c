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
