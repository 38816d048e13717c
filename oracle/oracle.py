"""TEST INFRASTRUCTURE ONLY -- Python face of the CPU oracle (oracle/quake_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this
module, and only as the checker or the reported baseline. Nothing under quake_b200/ imports it.
Every function names the reference code it restates; the C side carries the file:line citations.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

f32p = C.POINTER(C.c_float)
i64p = C.POINTER(C.c_int64)
i32p = C.POINTER(C.c_int32)
f64p = C.POINTER(C.c_double)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            subprocess.run(["make", "-s", "-C", _HERE], check=True)
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_l2sqr.restype = C.c_float
        _lib.orc_ip.restype = C.c_float
        _lib.orc_incomplete_beta.restype = C.c_double
        _lib.orc_incomplete_beta.argtypes = [C.c_double, C.c_double, C.c_double]
        _lib.orc_serial_scan.restype = C.c_int
        _lib.orc_recall_profile.restype = C.c_int
    return _lib


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _i64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int64))


def _p(a, t):
    return a.ctypes.data_as(t)


def _np(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def is_ip(metric) -> int:
    if isinstance(metric, str):
        return 1 if metric.lower() == "ip" else 0
    return 1 if int(metric) == 0 else 0  # faiss numbering: 0 = inner product, 1 = l2


def pairwise(x, y, metric="l2") -> np.ndarray:
    """fvec_L2sqr / fvec_inner_product for every pair (squared l2)."""
    x, y = _f32(_np(x)), _f32(_np(y))
    out = np.empty((x.shape[0], y.shape[0]), dtype=np.float32)
    lib().orc_pairwise(_p(x, f32p), C.c_int64(x.shape[0]), _p(y, f32p), C.c_int64(y.shape[0]), C.c_int64(x.shape[1]),
                       C.c_int(is_ip(metric)), _p(out, f32p))
    return out


def scan_list(q, vecs, ids, k, metric="l2"):
    """scan_list + TopkBuffer: one query x one list -> (ids, distances), min(k, n) entries, best first."""
    q, vecs = _f32(_np(q)), _f32(_np(vecs))
    n, d = (vecs.shape[0], q.shape[0])
    idp = None
    if ids is not None:
        ids = _i64(_np(ids))
        idp = _p(ids, i64p)
    oi = np.empty(k, dtype=np.int64)
    od = np.empty(k, dtype=np.float32)
    on = C.c_int(0)
    lib().orc_scan_list(_p(q, f32p), _p(vecs, f32p), idp, C.c_int64(n), C.c_int64(d), C.c_int(is_ip(metric)), C.c_int(k),
                        _p(oi, i64p), _p(od, f32p), C.byref(on))
    return oi[: on.value].copy(), od[: on.value].copy()


def topk_stream(dist, ids, k, descending, capacity=8192):
    """TypedTopKBuffer fed with a stream of (distance, id); returns (ids, distances, kth)."""
    dist, ids = _f32(_np(dist)), _i64(_np(ids))
    oi = np.empty(k, dtype=np.int64)
    od = np.empty(k, dtype=np.float32)
    on, kth = C.c_int(0), C.c_float(0)
    lib().orc_topk_stream(_p(dist, f32p), _p(ids, i64p), C.c_int64(dist.size), C.c_int(k), C.c_int(int(descending)),
                          C.c_int(capacity), _p(oi, i64p), _p(od, f32p), C.byref(on), C.byref(kth))
    return oi[: on.value].copy(), od[: on.value].copy(), kth.value


def batched_scan_list(queries, vecs, ids, k, metric="l2"):
    """batched_scan_list, exact (nx < 20) branch -> (ids [nq,k], dist [nq,k], counts [nq]), padded."""
    queries, vecs = _f32(_np(queries)), _f32(_np(vecs))
    nq, d = queries.shape
    n = vecs.shape[0]
    idp = None
    if ids is not None:
        ids = _i64(_np(ids))
        idp = _p(ids, i64p)
    oi = np.empty((nq, k), dtype=np.int64)
    od = np.empty((nq, k), dtype=np.float32)
    on = np.empty(nq, dtype=np.int32)
    lib().orc_batched_scan_list(_p(queries, f32p), C.c_int64(nq), _p(vecs, f32p), idp, C.c_int64(n), C.c_int64(d),
                                C.c_int(is_ip(metric)), C.c_int(k), _p(oi, i64p), _p(od, f32p), _p(on, i32p))
    return oi, od, on


def serial_scan(queries, lists, probe, k, metric="l2", recall_target=-1.0, recompute_threshold=0.001,
                use_precomputed=True, cand_centroids=None):
    """QueryCoordinator::serial_scan. lists: sequence of (vecs [n,d], ids [n]); probe [Q, nprobe] indices into
    `lists` (-1 = skip); cand_centroids [Q, nprobe, d] enables APS (rank-ordered candidate centroids).
    Returns (ids [Q,k], dist [Q,k], partitions_scanned [Q])."""
    queries = _f32(_np(queries))
    Q, d = queries.shape
    probe = _i64(_np(probe))
    nprobe = probe.shape[1]
    vec_arrs = [_f32(_np(v)).reshape(-1, d) for v, _ in lists]
    id_arrs = [_i64(_np(i)) for _, i in lists]
    L = len(lists)
    vp = (f32p * L)(*[_p(v, f32p) for v in vec_arrs])
    ip_ = (i64p * L)(*[_p(i, i64p) for i in id_arrs])
    ln = _i64([v.shape[0] for v in vec_arrs])
    cents_p = None
    keep = None
    if cand_centroids is not None and recall_target > 0:
        cc = _f32(_np(cand_centroids))
        keep = cc
        arr = (f32p * (Q * nprobe))()
        base = cc.ctypes.data
        for i in range(Q * nprobe):
            arr[i] = C.cast(base + i * d * 4, f32p)
        cents_p = arr
    oi = np.empty((Q, k), dtype=np.int64)
    od = np.empty((Q, k), dtype=np.float32)
    sc = np.zeros(Q, dtype=np.int32)
    rc = lib().orc_serial_scan(_p(queries, f32p), C.c_int64(Q), C.c_int64(d), vp, ip_, _p(ln, i64p), _p(probe, i64p),
                               C.c_int(nprobe), C.c_int(is_ip(metric)), C.c_int(k), C.c_float(recall_target),
                               C.c_float(recompute_threshold), C.c_int(int(use_precomputed)), cents_p, _p(oi, i64p),
                               _p(od, f32p), _p(sc, i32p))
    if rc != 0:
        raise RuntimeError("Boundary distances must have at least 2 partitions to create an estimate.")
    del keep
    return oi, od, sc


def coarse_topk(queries, centroids, centroid_ids, k, metric="l2", blas=None):
    """The parent (flat) index search = batched_scan_list over the centroid list: exact per-pair loop for
    fewer than 20 queries, else the BLAS form ||x||^2 + ||y||^2 - 2<x,y> clamped at 0
    (faiss/utils/distances.cpp:262-343, threshold :649). Returns (ids [Q,k], dist [Q,k])."""
    queries, centroids = _f32(_np(queries)), _f32(_np(centroids))
    centroid_ids = _i64(_np(centroid_ids))
    Q = queries.shape[0]
    k = min(k, centroids.shape[0])
    if blas is None:
        blas = Q >= 20
    if not blas:
        oi, od, _ = batched_scan_list(queries, centroids, centroid_ids, k, metric)
        return oi, od
    ip = is_ip(metric)
    dots = queries @ centroids.T
    if ip:
        score = -dots
    else:
        xn = np.array([lib().orc_ip(_p(q, f32p), _p(q, f32p), C.c_size_t(q.size)) for q in queries], dtype=np.float32)
        yn = np.array([lib().orc_ip(_p(c, f32p), _p(c, f32p), C.c_size_t(c.size)) for c in centroids], dtype=np.float32)
        score = (xn[:, None] + yn[None, :]) - np.float32(2) * dots
        score = np.maximum(score, np.float32(0))
    order = np.lexsort((np.broadcast_to(np.arange(score.shape[1]), score.shape), score), axis=1)[:, :k]
    sel = np.take_along_axis(score, order, axis=1)
    dist = -sel if ip else np.sqrt(sel)
    return centroid_ids[order], dist.astype(np.float32)


def assign(x, centroids, metric="l2") -> np.ndarray:
    x, c = _f32(_np(x)), _f32(_np(centroids))
    out = np.empty(x.shape[0], dtype=np.int32)
    lib().orc_assign(_p(x, f32p), C.c_int64(x.shape[0]), _p(c, f32p), C.c_int64(c.shape[0]), C.c_int64(x.shape[1]),
                     C.c_int(is_ip(metric)), _p(out, i32p))
    return out


def centroid_sums(x, assign_, K):
    x = _f32(_np(x))
    a = np.ascontiguousarray(np.asarray(_np(assign_), dtype=np.int32))
    sums = np.empty((K, x.shape[1]), dtype=np.float32)
    counts = np.empty(K, dtype=np.int64)
    lib().orc_centroid_sums(_p(x, f32p), C.c_int64(x.shape[0]), C.c_int64(x.shape[1]), _p(a, i32p), C.c_int64(K),
                            _p(sums, f32p), _p(counts, i64p))
    return sums, counts


def kmeans_refine(centroids, parts, metric="l2", iterations=0):
    """kmeans_refine_partitions (clustering.cpp:99-182). parts: list of (vecs, ids). Returns
    (centroids used for the last assignment, new parts)."""
    centroids = _f32(_np(centroids)).copy()
    K, d = centroids.shape
    parts = [(_f32(_np(v)).reshape(-1, d), _i64(_np(i))) for v, i in parts]
    iters = iterations if iterations > 0 else 1
    sums = np.zeros((K, d), np.float32)
    counts = np.zeros(K, np.int64)
    for it in range(iters):
        if it > 0:
            with np.errstate(invalid="ignore", divide="ignore"):
                centroids = (sums / counts[:, None].astype(np.float32)).astype(np.float32)
        allv = np.concatenate([v for v, _ in parts]) if parts else np.zeros((0, d), np.float32)
        alli = np.concatenate([i for _, i in parts]) if parts else np.zeros((0,), np.int64)
        if allv.shape[0]:
            a = assign(allv, centroids, metric)
        else:
            a = np.zeros(0, np.int32)
        sums, counts = centroid_sums(allv, a, K)
        parts = [(allv[a == c], alli[a == c]) for c in range(K)]
    return centroids, parts


def boundary_distances(q, cents, euclid=True) -> np.ndarray:
    q, cents = _f32(_np(q)), _f32(_np(cents))
    m, d = cents.shape
    arr = (f32p * m)(*[C.cast(cents.ctypes.data + j * d * 4, f32p) for j in range(m)])
    out = np.empty(m, dtype=np.float32)
    lib().orc_boundary_distances(_p(q, f32p), arr, C.c_int(m), C.c_int(d), C.c_int(int(euclid)), _p(out, f32p))
    return out


def recall_profile(boundary, radius, d, use_precomputed=True, euclid=True) -> np.ndarray:
    b = _f32(boundary)
    table = np.empty(1001, dtype=np.float64)
    lib().orc_beta_table(C.c_int(d), _p(table, f64p))
    out = np.empty(b.size, dtype=np.float32)
    rc = lib().orc_recall_profile(_p(b, f32p), C.c_int(b.size), C.c_float(radius), C.c_int(d), C.c_int(int(use_precomputed)),
                                  C.c_int(int(euclid)), _p(table, f64p), _p(out, f32p))
    if rc:
        raise RuntimeError("Boundary distances must have at least 2 partitions to create an estimate.")
    return out


def incomplete_beta(a, b, x) -> float:
    return lib().orc_incomplete_beta(a, b, x)


# ------------------------------------------------------------------------------------------------
# helpers that read an index OBJECT (ours or the compiled reference's save files) for parity runs
# ------------------------------------------------------------------------------------------------
def index_lists(idx):
    """(partition ids, [(vecs, ids)], centroids, centroid ids) of a quake_b200.QuakeIndex, on the CPU."""
    pids = idx.store.partition_ids()
    lists = []
    for p in pids:
        v, i = idx.store.get_list(int(p))
        lists.append((v.cpu().numpy().copy(), i.cpu().numpy().copy()))
    if idx.parent is not None:
        cv, ci = idx.parent.store.get_list(0)
        return pids, lists, cv.cpu().numpy().copy(), ci.cpu().numpy().copy()
    return pids, lists, None, None


def read_index_dir(path):
    """Parse an index directory in the reference's on-disk format (quake_index.cpp:170-267,
    dynamic_inverted_list.cpp:338-419) -> (metric, pids, [(vecs, ids)], centroids | None, centroid ids | None)."""
    import struct

    def read_partitions(fn):
        raw = open(fn, "rb").read()
        magic, version, _nl, code_size, nparts = struct.unpack_from("<IIQQQ", raw, 0)
        assert magic == 0x44494E4C and version == 3
        d = code_size // 4
        offs = np.frombuffer(raw, np.uint64, nparts + 1, 32)
        pids = np.frombuffer(raw, np.uint64, nparts, 32 + 8 * (nparts + 1)).astype(np.int64)
        start = 32 + 8 * (nparts + 1) + 8 * nparts
        lists = []
        for i in range(nparts):
            nv = int(offs[i + 1] - offs[i]) // (code_size + 8)
            o = start + int(offs[i])
            v = np.frombuffer(raw, np.float32, nv * d, o).reshape(nv, d).copy()
            ids = np.frombuffer(raw, np.int64, nv, o + nv * code_size).copy()
            lists.append((v, ids))
        return pids, lists

    metric = 1
    for line in open(os.path.join(path, "metadata.txt")):
        if line.startswith("metric="):
            metric = int(line.strip().split("=")[1])
    pids, lists = read_partitions(os.path.join(path, "partitions"))
    order = np.argsort(pids)
    pids, lists = pids[order], [lists[i] for i in order]
    cv = ci = None
    pp = os.path.join(path, "parent", "partitions")
    if os.path.exists(pp):
        _, pl = read_partitions(pp)
        cv, ci = pl[0]
    return metric, pids, lists, cv, ci


def search_lists(pids, lists, cv, ci, q, k, nprobe, metric, recall_target=-1.0, initial_search_fraction=0.02,
                 recompute_threshold=0.001, use_precomputed=True, blas=None, return_probe=False):
    """QueryCoordinator::search (query_coordinator.cpp:612-657) restated over explicit lists: coarse scan of
    the centroid list (cv, ci), then serial_scan of the probed lists. `blas` selects the coarse-scan
    arithmetic (None = the reference's rule: sgemm form for >= 20 queries). Returns torch (ids, distances)."""
    qn = _f32(_np(q))
    Q = qn.shape[0]
    if cv is None:
        probe = np.tile(np.arange(len(lists), dtype=np.int64), (Q, 1))
        oi, od, _ = serial_scan(qn, lists, probe, k, metric)
        return torch.from_numpy(oi), torch.from_numpy(od)
    nlist = len(lists)
    use_aps = recall_target > 0
    kp = max(int(nlist * initial_search_fraction), 1) if use_aps else min(nprobe, nlist)
    cid, _ = coarse_topk(qn, cv, ci, kp, metric, blas=blas)
    slot_of = {int(p): s for s, p in enumerate(pids)}
    probe = np.vectorize(lambda p: slot_of.get(int(p), -1))(cid).astype(np.int64)
    cents = None
    if use_aps:
        row_of = {int(c): r for r, c in enumerate(ci)}
        rows = np.vectorize(lambda p: row_of[int(p)])(cid)
        cents = cv[rows]
    oi, od, sc = serial_scan(qn, lists, probe, k, metric, recall_target, recompute_threshold, use_precomputed, cents)
    if return_probe:
        return torch.from_numpy(oi), torch.from_numpy(od), cid, sc
    return torch.from_numpy(oi), torch.from_numpy(od)


def search_index_like(idx, q, k, nprobe, metric=None, **kw):
    """search_lists over the CONTENT of a quake_b200.QuakeIndex (same centroids + list membership)."""
    metric = idx.metric if metric is None else metric
    pids, lists, cv, ci = index_lists(idx)
    return search_lists(pids, lists, cv, ci, q, k, nprobe, metric, **kw)
