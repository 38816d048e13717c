"""Pins the CPU oracle (oracle/quake_oracle.c + oracle/oracle.py) against golden vectors generated from
the compiled, unmodified reference (tests/golden/make_golden.py), against the reference's own
known-answer tests, and -- where oracle/_ref is importable -- against the live reference."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as orc
from tests.conftest import GOLDEN

U = np.load(os.path.join(GOLDEN, "unit.npz"))
S = np.load(os.path.join(GOLDEN, "search.npz"))


@pytest.mark.parametrize("d", [3, 8, 12, 13, 32, 96, 100, 128, 131])
def test_pairwise_bit_exact(d):
    """fvec_L2sqr / fvec_inner_product (faiss/utils/distances_simd.cpp:188-224): bit-identical."""
    x, y = U[f"pw_x_{d}"], U[f"pw_y_{d}"]
    assert np.array_equal(orc.pairwise(x, y, "l2"), U[f"pw_l2_{d}"])
    assert np.array_equal(orc.pairwise(x, y, "ip"), U[f"pw_ip_{d}"])


@pytest.mark.parametrize("name", ["a", "b", "c", "d"])
@pytest.mark.parametrize("metric", ["l2", "ip"])
def test_scan_list_golden(name, metric):
    vecs, ids, qs, k = U[f"sl_{name}_vecs"], U[f"sl_{name}_ids"], U[f"sl_{name}_q"], int(U[f"sl_{name}_k"])
    for i in range(qs.shape[0]):
        oi, od = orc.scan_list(qs[i], vecs, ids, k, metric)
        n = oi.size
        assert np.array_equal(oi, U[f"sl_{name}_{metric}_ids"][i, :n])
        assert np.array_equal(od, U[f"sl_{name}_{metric}_dist"][i, :n])
        assert (U[f"sl_{name}_{metric}_ids"][i, n:] == -1).all()
    bi, bd, bc = orc.batched_scan_list(qs, vecs, ids, k, metric)
    assert np.array_equal(bc, U[f"bsl_{name}_{metric}_cnt"])
    assert np.array_equal(bi, U[f"bsl_{name}_{metric}_ids"])
    assert np.array_equal(bd, U[f"bsl_{name}_{metric}_dist"])


@pytest.mark.parametrize("name", ["a", "b", "c"])
@pytest.mark.parametrize("desc", [0, 1])
def test_topk_buffer_golden(name, desc):
    """TypedTopKBuffer (list_scanning.h:41-204), incl. overflow flushes (capacity < stream length)."""
    k, cap = U[f"tk_{name}_kcap"]
    i, d, kth = orc.topk_stream(U[f"tk_{name}_dist"], U[f"tk_{name}_ids"], int(k), desc, int(cap))
    assert np.array_equal(i, U[f"tk_{name}_{desc}_ids"])
    assert np.array_equal(d, U[f"tk_{name}_{desc}_d"])
    assert np.float32(kth) == U[f"tk_{name}_{desc}_kth"]


def test_known_answer_scan_list_l2_ip():
    """test/cpp/list_scanning.cpp:22-55, 58-91 (hand-written vectors; ties may resolve either way)."""
    q = np.array([1, 0, 0], np.float32)
    v = np.array([[1, 0, 0], [0, 1, 0], [1, 1, 0], [2, 0, 0]], np.float32)
    ids = np.array([10, 20, 30, 40])
    oi, od = orc.scan_list(q, v, ids, 2, "l2")
    assert od.tolist() == [0.0, 1.0] and oi[0] == 10 and oi[1] in (30, 40)
    oi, od = orc.scan_list(q, v, ids, 2, "ip")
    assert od.tolist() == [2.0, 1.0] and oi[0] == 40 and oi[1] in (10, 30)


def test_known_answer_batched_and_edge_cases():
    """test/cpp/list_scanning.cpp:94-149 (batched l2), :281-341 (no ids => local offsets),
    :344-382 (empty list), :386-430 (list shorter than k; second distance sqrt(3))."""
    qs = np.array([[1, 0, 0], [0, 1, 0]], np.float32)
    v = np.array([[1, 0, 0], [0, 1, 0], [1, 1, 0], [2, 0, 0]], np.float32)
    ids = np.array([10, 20, 30, 40])
    bi, bd, bc = orc.batched_scan_list(qs, v, ids, 2, "l2")
    assert bi.tolist() == [[10, 30], [20, 30]] and bd.tolist() == [[0.0, 1.0], [0.0, 1.0]]
    bi, bd, bc = orc.batched_scan_list(qs, v, None, 2, "l2")
    assert bi.tolist() == [[0, 2], [1, 2]]
    bi, bd, bc = orc.batched_scan_list(qs, np.zeros((0, 3), np.float32), None, 2, "l2")
    assert bc.tolist() == [0, 0] and (bi == -1).all()
    q1 = np.array([[1, 1, 1]], np.float32)
    v2 = np.array([[1, 1, 1], [2, 2, 2]], np.float32)
    bi, bd, bc = orc.batched_scan_list(q1, v2, np.array([100, 200]), 5, "l2")
    assert bc.tolist() == [2] and bi[0, :2].tolist() == [100, 200]
    assert bd[0, 0] == 0.0 and bd[0, 1] == np.float32(np.sqrt(np.float32(3.0)))


def test_scan_list_equals_torch_topk():
    """test/cpp/list_scanning.cpp:432-496: scan_list top-10 ids == torch.topk(cdist / matmul) ids."""
    g = torch.Generator().manual_seed(5)
    x = torch.randn(3000, 128, generator=g)
    qs = torch.randn(20, 128, generator=g)
    ids = torch.arange(3000)
    for i in range(20):
        oi, od = orc.scan_list(qs[i], x, ids, 10, "l2")
        td, ti = torch.cdist(qs[i:i + 1], x).topk(10, largest=False)
        assert oi.tolist() == ti[0].tolist() and np.allclose(od, td[0].numpy(), atol=1e-2)
        oi, od = orc.scan_list(qs[i], x, ids, 10, "ip")
        td, ti = (qs[i:i + 1] @ x.T).topk(10)
        assert oi.tolist() == ti[0].tolist()


@pytest.mark.parametrize("d", [16, 128])
def test_aps_geometry_golden(d):
    """geometry.h: incomplete_beta (:115-161), compute_boundary_distances (:57-113), compute_recall_profile
    (:345-407)."""
    a, b = (d + 1) / 2.0, 0.5
    mine = np.array([orc.incomplete_beta(a, b, float(x)) for x in U["beta_x"]])
    assert np.allclose(mine, U[f"beta_{d}"], rtol=1e-12, atol=1e-300)
    q, c = U[f"bd_q_{d}"], U[f"bd_c_{d}"]
    bd = orc.boundary_distances(q, c, True)
    assert np.allclose(bd, U[f"bd_l2_{d}"], rtol=2e-6)
    qn, cn = q / np.linalg.norm(q), c / np.linalg.norm(c, axis=1, keepdims=True)
    assert np.allclose(orc.boundary_distances(qn.astype(np.float32), cn.astype(np.float32), False), U[f"bd_ip_{d}"],
                       rtol=1e-4, atol=1e-5)
    for j, r in enumerate(U[f"rp_radii_{d}"]):
        for pre in (0, 1):
            key = f"rp_{d}_{j}_{pre}"
            if key not in U:
                continue
            mine = orc.recall_profile(U[f"bd_l2_{d}"], float(r), d, bool(pre), True)
            assert np.allclose(mine, U[key], rtol=1e-5, atol=1e-7), key


@pytest.mark.parametrize("metric", ["l2", "ip"])
@pytest.mark.parametrize("iters", [0, 3])
def test_kmeans_refine_golden(metric, iters):
    """kmeans_refine_partitions (clustering.cpp:99-182): same list membership (ids, in order), centroids
    within 1e-5 relative."""
    cents, sizes, vecs = U[f"rf_{metric}_cents"], U[f"rf_{metric}_sizes"], U[f"rf_{metric}_vecs"]
    offs = np.concatenate([[0], np.cumsum(sizes)])
    parts = [(vecs[offs[i]:offs[i + 1]], np.arange(offs[i], offs[i + 1])) for i in range(len(sizes))]
    c2, newp = orc.kmeans_refine(cents, parts, metric, iters)
    assert np.array_equal(np.array([p[0].shape[0] for p in newp]), U[f"rf_{metric}_{iters}_sizes"])
    assert np.array_equal(np.concatenate([p[1] for p in newp]), U[f"rf_{metric}_{iters}_ids"])
    ref = U[f"rf_{metric}_{iters}_cents"]
    assert np.array_equal(np.isnan(c2), np.isnan(ref))
    ok = ~np.isnan(ref)
    assert np.allclose(c2[ok], ref[ok], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("metric", ["l2", "ip"])
@pytest.mark.parametrize("tag", ["serial_small", "serial", "batched", "k100", "all"])
def test_search_golden(metric, tag):
    """QuakeIndex.search of the reference on an index the reference built and saved: the oracle, reading the
    same files, returns the same ids; distances bit-identical on the serial path, within 1e-4 relative on the
    batched (sgemm) path."""
    m, pids, lists, cv, ci = orc.read_index_dir(os.path.join(GOLDEN, f"index_{metric}"))
    nq, k, nprobe, batched = S[f"{metric}_{tag}_cfg"]
    q = S[f"{metric}_q"][:nq]
    oi, od = orc.search_lists(pids, lists, cv, ci, q, int(k), int(nprobe), m)
    rid, rd = S[f"{metric}_{tag}_ids"], S[f"{metric}_{tag}_dist"]
    assert np.array_equal(oi.numpy(), rid)
    if batched:
        assert np.allclose(od.numpy(), rd, rtol=1e-4)
    else:
        assert np.array_equal(od.numpy(), rd)


@pytest.mark.parametrize("metric", ["l2", "ip"])
def test_search_aps_golden(metric):
    """APS (recall_target=0.9) through serial_scan (query_coordinator.cpp:521-579)."""
    m, pids, lists, cv, ci = orc.read_index_dir(os.path.join(GOLDEN, f"index_{metric}"))
    q = S[f"{metric}_q"]
    oi, od = orc.search_lists(pids, lists, cv, ci, q, 10, 0, m, recall_target=0.9, initial_search_fraction=0.1,
                              use_precomputed=False)
    assert np.array_equal(oi.numpy(), S[f"{metric}_aps_ids"])
    assert np.array_equal(od.numpy(), S[f"{metric}_aps_dist"])


def test_against_live_reference(quake_ref):
    """Where oracle/_ref is importable: fresh random inputs through both."""
    g = torch.Generator().manual_seed(77)
    x, y = torch.randn(11, 72, generator=g), torch.randn(200, 72, generator=g)
    for m in ("l2", "ip"):
        assert np.array_equal(quake_ref.shim.pairwise(x, y, m).numpy(), orc.pairwise(x, y, m))
        ids = torch.arange(200) * 2
        a, b = quake_ref.shim.scan_list(x[0], y, ids, 7, m)
        oi, od = orc.scan_list(x[0], y, ids, 7, m)
        assert a.tolist() == oi.tolist() and np.array_equal(b.numpy(), od)


def test_avx512_golden_set_has_the_same_ids():
    """tests/golden/search_avx512.npz comes from an AVX-512 build of the unmodified reference (the ISA its own
    -march=native build uses on the survey host), search.npz from the AVX2 build the refine kernel reproduces bit for
    bit: identical ids, distances equal to ~1e-7 relative (16-lane instead of 8-lane summation)."""
    A = np.load(os.path.join(GOLDEN, "search.npz"))
    B = np.load(os.path.join(GOLDEN, "search_avx512.npz"))
    checked = 0
    for key in B.files:
        if key.startswith("d128") or key.endswith("_cfg") or key.endswith("_q"):
            continue
        a, b = A[key], B[key]
        if key.endswith("_ids"):
            assert np.array_equal(a, b), key
        else:
            fin = np.isfinite(a)
            assert np.array_equal(a[~fin], b[~fin])
            assert np.allclose(a[fin], b[fin], rtol=1e-6, atol=0), key
        checked += 1
    assert checked == 20
