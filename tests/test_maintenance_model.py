"""The maintenance cost model (quake_b200/maintenance.py) against the compiled reference's MaintenanceCostEstimator /
ListScanLatencyEstimator (src/cpp/src/maintenance_cost_estimator.cpp:131-251, 355-493) on the SAME latency table
(oracle/_ref shim: CostModel). Host logic only -- no GPU."""
import numpy as np
import pytest

from quake_b200 import maintenance as mt


def _table():
    """A convex, k-dependent synthetic latency grid [n_values x k_values] in ns."""
    n = np.array(mt.LATENCY_RANGE_N, dtype=np.float64)[:, None]
    k = np.array(mt.LATENCY_RANGE_K, dtype=np.float64)[None, :]
    return (400.0 + 1.7 * n + 0.00002 * n * n + 35.0 * np.log2(k + 1.0) + 0.01 * n * np.log2(k + 1.0)).astype(np.float32)


def test_interpolation_matches_reference(quake_ref):
    t = _table()
    ref = quake_ref.shim.CostModel(64, 0.9, 10, t.tolist())
    assert ref.n_values() == mt.LATENCY_RANGE_N and ref.k_values() == mt.LATENCY_RANGE_K
    lat = mt.ListScanLatencyEstimator(64, table=t)
    for n in (1, 2, 3, 5, 16, 17, 100, 256, 1000, 4096, 50000, 65536, 70000, 200000):
        for k in (1, 2, 4, 10, 16, 100, 256, 300, 1000):
            want = ref.latency(n, k)
            got = lat.estimate_scan_latency(n, k)
            assert got == pytest.approx(want, rel=1e-5, abs=1e-3), (n, k)
    assert lat.estimate_scan_latency(0, 10) == 0.0


def test_deltas_match_reference(quake_ref):
    t = _table()
    ref = quake_ref.shim.CostModel(64, 0.9, 10, t.tolist())
    est = mt.MaintenanceCostEstimator(64, 0.9, 10, mt.ListScanLatencyEstimator(64, table=t))
    rng = np.random.default_rng(5)
    for _ in range(200):
        size = int(rng.integers(1, 5000))
        rate = float(rng.random()) * 0.3
        total = int(rng.integers(2, 20000))
        avg_rate = float(rng.random()) * 0.05
        avg_size = int(rng.integers(1, 3000))
        assert est.compute_split_delta(size, rate, total) == pytest.approx(ref.split_delta(size, rate, total), rel=1e-4, abs=1e-2)
        # cost_old / cost_new are ~T * rate * L each (up to 1e6 here) in single precision, and the compiled reference
        # contracts a*b+c into FMAs: agreement to a few ulps of those intermediates
        ulp = 1.2e-7 * total * max(avg_rate, rate) * float(t.max())
        assert est.compute_delete_delta(size, rate, total, avg_rate, avg_size) == pytest.approx(
            ref.delete_delta(size, rate, total, avg_rate, avg_size), rel=2e-4, abs=5e-2 + 8 * ulp)
        m = int(rng.integers(1, 6))
        counts = [int(x) for x in rng.integers(1, 50, m)]
        sizes = [int(x) for x in rng.integers(1, 4000, m)]
        rates = [float(x) * 0.2 for x in rng.random(m)]
        assert est.compute_delete_delta_w_reassign(size, rate, total, counts, sizes, rates) == pytest.approx(
            ref.delete_delta_w_reassign(size, rate, total, counts, sizes, rates), rel=2e-4, abs=5e-2)
    assert est.compute_delete_delta(100, 0.1, 1, 0.1, 100) == 0.0 == ref.delete_delta(100, 0.1, 1, 0.1, 100)


def test_hit_window_semantics():
    """HitCountTracker (hit_count_tracker.cpp:43-66) for batches: only the last window_size queries count."""
    import torch
    tr = mt.HitCountTracker(5, 100)
    tr.add_batch(torch.tensor([[0, 1], [1, 2], [2, 3]]), None)
    assert tr.num_queries_recorded == 3
    tr.add_batch(torch.tensor([[4, 4, -1], [5, 6, 7], [0, -1, -1], [9, 9, 9]]), torch.tensor([2, 3, 1, 0]))
    assert tr.num_queries_recorded == 5
    sizes = torch.arange(10, dtype=torch.int64) * 10
    ids, nq, frac = tr.window(sizes)
    assert nq == 5
    assert ids.tolist() == [[2, 3, -1], [4, 4, -1], [5, 6, 7], [0, -1, -1], [-1, -1, -1]]
    want = np.mean([(20 + 30) / 100, (40 + 40) / 100, (50 + 60 + 70) / 100, 0.0, 0.0])
    assert frac == pytest.approx(want)
