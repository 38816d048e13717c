"""GPU parity of the partition scan (qk_scan_partitions, through the C ABI) against the CPU oracle.

Bar (BASELINE.json north_star): returned ids bit-identical to the reference CPU path at fixed nprobe;
distances within 1e-4 relative (the kernel's refine step actually reproduces the reference's summation
order, so they are compared for exact equality first and the count of inexact ones is reported).
"""
import numpy as np
import pytest
import torch

from oracle import oracle as orc

pytestmark = pytest.mark.gpu

REL_TOL = 1e-4  # north_star: L2/IP distances within 1e-4 relative


def _make_store(sizes, d, seed, metric="l2", dup=False):
    import quake_b200 as qb
    from quake_b200.store import PartitionStore
    from quake_b200 import clustering

    g = torch.Generator().manual_seed(seed)
    n = int(sum(sizes))
    x = torch.randn(n, d, generator=g)
    if dup and n > 8:
        x[n // 2:] = x[: n - n // 2]  # exact duplicates => exact distance ties
    ids = torch.randperm(n, generator=g).to(torch.int64) + 7
    dev = torch.device("cuda", 0)
    st = PartitionStore(d, dev)
    xd = clustering.pad_rows(x, dev)
    st.init_from_sorted(xd, ids.to(dev), None, np.asarray(sizes, dtype=np.int64), np.arange(len(sizes), dtype=np.int64))
    offs = np.concatenate([[0], np.cumsum(sizes)])
    lists = [(x[offs[i]:offs[i + 1]].numpy(), ids[offs[i]:offs[i + 1]].numpy()) for i in range(len(sizes))]
    return st, lists


def _check(st, lists, Q, nprobe, k, metric, seed, skip_frac=0.0):
    from quake_b200 import index as qidx, clustering, _lib

    d = st.d
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(Q, d, generator=g)
    L = len(lists)
    probe = torch.stack([torch.randperm(L, generator=g)[:nprobe] for _ in range(Q)]).to(torch.int64)
    if skip_frac > 0:
        mask = torch.rand(probe.shape, generator=g) < skip_frac
        probe[mask] = -1
    m = _lib.QK_METRIC_INNER_PRODUCT if metric == "ip" else _lib.QK_METRIC_L2
    xq = clustering.pad_rows(q, st.device)
    ids, dist = qidx.scan_partitions(st, xq, probe.to(torch.int32).to(st.device), k, m)
    torch.cuda.synchronize()
    oi, od, _ = orc.serial_scan(q.numpy(), lists, probe.numpy(), k, metric)
    ids, dist = ids.cpu().numpy(), dist.cpu().numpy()
    assert np.array_equal(ids, oi), f"ids differ in {(ids != oi).sum()} of {ids.size} slots"
    fin = np.isfinite(od)
    assert np.array_equal(np.isfinite(dist), fin)
    assert np.array_equal(dist[~fin], od[~fin])  # +-inf padding
    rel = np.abs(dist[fin] - od[fin]) / np.maximum(np.abs(od[fin]), 1e-30)
    assert rel.size == 0 or rel.max() <= REL_TOL
    return int((dist[fin] != od[fin]).sum())


@pytest.mark.parametrize("metric", ["l2", "ip"])
@pytest.mark.parametrize("d", [128, 96, 32, 100, 3, 200])
def test_scan_matches_oracle(metric, d):
    g = np.random.default_rng(d)
    sizes = g.integers(0, 700, size=64)
    sizes[3] = 0
    sizes[5] = 1
    st, lists = _make_store(sizes, d, seed=d)
    inexact = _check(st, lists, Q=200, nprobe=8, k=10, metric=metric, seed=1)
    assert inexact == 0, f"{inexact} distances not bit-identical to the oracle"


@pytest.mark.parametrize("k", [1, 10, 26, 27, 100, 128, 200])
def test_scan_k_sweep(k):
    sizes = np.full(32, 300)
    st, lists = _make_store(sizes, 64, seed=k)
    _check(st, lists, Q=70, nprobe=6, k=k, metric="l2", seed=2)
    _check(st, lists, Q=70, nprobe=6, k=k, metric="ip", seed=3)


def test_scan_fewer_than_k_and_skips():
    """k > available => id -1 and +inf (l2) / -inf (ip) (query_coordinator.cpp:589-601); -1 probes skipped."""
    sizes = np.array([0, 2, 3, 0, 1, 4])
    st, lists = _make_store(sizes, 16, seed=5)
    _check(st, lists, Q=33, nprobe=3, k=10, metric="l2", seed=4, skip_frac=0.3)
    _check(st, lists, Q=33, nprobe=3, k=10, metric="ip", seed=4, skip_frac=0.3)


def test_scan_large_list_segments():
    """A list longer than QK_SEGMENT_ROWS is cut into several scan segments (flat index case)."""
    sizes = np.array([10000, 5000, 4096, 4097])
    st, lists = _make_store(sizes, 128, seed=6)
    _check(st, lists, Q=40, nprobe=4, k=10, metric="l2", seed=5)
    _check(st, lists, Q=40, nprobe=2, k=100, metric="ip", seed=6)


def test_scan_duplicates_tie_break():
    """Exact duplicate vectors give exact distance ties: the order is (distance, id), one valid outcome
    of the reference's distance-only comparator, and the one the oracle fixes."""
    sizes = np.full(8, 200)
    st, lists = _make_store(sizes, 32, seed=7, dup=True)
    _check(st, lists, Q=50, nprobe=8, k=10, metric="l2", seed=7)
    _check(st, lists, Q=50, nprobe=8, k=10, metric="ip", seed=8)


def test_forced_exact_rescan_path(monkeypatch):
    """The exhaustive exact re-scan (taken when the filter's error bound cannot prove the top-k) must give
    the same answer as the filter+refine path."""
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r);"
            "import numpy as np, tests.test_gpu_scan as t;"
            "st, lists = t._make_store(np.full(16, 257), 128, seed=9);"
            "t._check(st, lists, Q=64, nprobe=5, k=10, metric='l2', seed=9);"
            "t._check(st, lists, Q=64, nprobe=5, k=10, metric='ip', seed=9); print('ok')") % (root, os.path.join(root, 'tests'))
    env = dict(os.environ, QK_FORCE_RESCAN="1")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, cwd=root)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


def test_fp32_pipe_kernel_path():
    """d <= 128 normally runs the tensor-core filter; QK_SCAN_PATH=ffma forces the FP32-pipe kernel (the one
    larger d uses) over the same cases."""
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys; sys.path.insert(0, %r);"
            "import numpy as np, tests.test_gpu_scan as t;"
            "g = np.random.default_rng(1); sizes = g.integers(0, 700, size=64); sizes[3] = 0; sizes[5] = 1;"
            "st, lists = t._make_store(sizes, 128, seed=3);"
            "assert t._check(st, lists, Q=200, nprobe=8, k=10, metric='l2', seed=1) == 0;"
            "assert t._check(st, lists, Q=200, nprobe=8, k=100, metric='ip', seed=2) == 0;"
            "st, lists = t._make_store(sizes, 100, seed=4);"
            "assert t._check(st, lists, Q=70, nprobe=8, k=10, metric='l2', seed=1) == 0; print('ok')") % (root,)
    env = dict(os.environ, QK_SCAN_PATH="ffma")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, cwd=root)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


def test_large_k_many_probes_no_rescan():
    """k = 100 over 32 probed lists of ~600 rows: thresholds converge slowly for large k, the candidate buffers must
    absorb that without sending queries to the exhaustive re-scan (and the answer is still the oracle's)."""
    import os
    from quake_b200 import index as qidx
    sizes = np.full(48, 600)
    st, lists = _make_store(sizes, 128, seed=21)
    os.environ["QK_SCAN_STATS"] = "1"
    try:
        _check(st, lists, Q=256, nprobe=32, k=100, metric="l2", seed=22)
        stats = qidx.LAST_SCAN_STATS.cpu().tolist()
    finally:
        os.environ.pop("QK_SCAN_STATS")
    assert stats[0] == 0, f"{stats[0]} queries took the exact re-scan"


def test_flat_store_threshold_path_matches_dense_path():
    """A single-list store is scanned in dense mode by default (every key stored, one select per query);
    QK_NO_DENSE=1 forces the threshold/seed path over the same cases."""
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys; sys.path.insert(0, %r);"
            "import numpy as np, tests.test_gpu_scan as t;"
            "st, lists = t._make_store(np.array([5000]), 128, seed=31);"
            "assert t._check(st, lists, Q=300, nprobe=1, k=64, metric='l2', seed=1) == 0;"
            "assert t._check(st, lists, Q=300, nprobe=1, k=10, metric='ip', seed=2) == 0; print('ok')") % (root,)
    for env_extra in ({}, {"QK_NO_DENSE": "1"}):
        env = dict(os.environ, **env_extra)
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, cwd=root)
        assert out.returncode == 0 and "ok" in out.stdout, (env_extra, out.stderr[-2000:])
