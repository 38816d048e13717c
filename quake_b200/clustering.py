"""k-means build / refit loop on the device.

Host-side mirror of the reference's ``kmeans`` and ``kmeans_refine_partitions``
(/root/reference/src/cpp/src/clustering.cpp:13-97, 99-182) and of the ``faiss::Clustering::train``
loop they delegate to (third_party/faiss/faiss/Clustering.cpp:255-539). The arithmetic runs in the
CUDA kernels behind the C ABI (assign: qk_kmeans_assign; update: qk_partition_by_assignment +
qk_kmeans_accumulate); this file is the control flow only.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr

MAX_POINTS_PER_CENTROID = 256  # faiss::ClusteringParameters default
FAISS_SEED = 1234


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def pad_rows(x: torch.Tensor, device) -> torch.Tensor:
    """float32 [n, pitch] copy on `device`, pitch = d rounded up to 4, zero padded."""
    n, d = x.shape
    pitch = (d + 3) // 4 * 4
    x = x.to(device=device, dtype=torch.float32)
    if pitch == d:
        return x.contiguous()
    out = torch.zeros((n, pitch), dtype=torch.float32, device=device)
    out[:, :d] = x
    return out


def rand_perm_prefix(n: int, seed: int, m: int) -> np.ndarray:
    lib = _lib.load()
    out = np.empty(m, dtype=np.int64)
    check(lib.qk_host_rand_perm_prefix(n, seed, m, out.ctypes.data_as(C.POINTER(C.c_int64))))
    return out


ASSIGN_FILTER_POLICY = "auto"  # "auto": 2xTF32 until the evidence says otherwise; "3": always 3xTF32; "2": always 2xTF32
_ASSIGN_RESCAN_LIMIT = 0.005   # fraction of a call's points sent to the exact re-scan before a run falls back to 3 terms


class AssignFilter:
    """Filter precision of the k-means assign over one training / refit run (include/quake_b200.h:
    qk_kmeans_assign_filtered). The assignment itself is exact either way; this only decides how fast it is found."""

    def __init__(self, device):
        import os
        policy = os.environ.get("QK_ASSIGN_FILTER", ASSIGN_FILTER_POLICY)
        self.fixed = policy in ("2", "3")
        self.terms = 3 if policy == "3" else 2
        self.stats = torch.zeros(2, dtype=torch.int32, device=device)
        self.points = 0

    def review(self) -> None:
        """Reads the evidence of the calls since the last review (one small D2H copy: call it where the host
        synchronises anyway) and falls back to 3 terms for good when too many points were re-scanned."""
        if self.fixed or self.terms == 3 or self.points == 0:
            return
        rescanned = int(self.stats[0].item())
        if rescanned > max(8, int(_ASSIGN_RESCAN_LIMIT * self.points)):
            self.terms = 3
        self.stats.zero_()
        self.points = 0


def assign_points(x: torch.Tensor, d: int, centroids: torch.Tensor, metric: int, want_dist: bool = False,
                  filt: AssignFilter | None = None):
    """Nearest centroid of every row of x ([n, pitch] device) -> int32 [n] (and distances). `filt`: the run's
    filter-precision state (None: 3xTF32)."""
    lib = _lib.load()
    _lib.require_device()
    n, K = int(x.shape[0]), int(centroids.shape[0])
    out = torch.empty(n, dtype=torch.int32, device=x.device)
    dist = torch.empty(n, dtype=torch.float32, device=x.device) if want_dist else None
    wsb = lib.qk_kmeans_assign_workspace_bytes(n, K, d)
    if wsb == 0:
        check(1)
    ws = torch.empty(wsb, dtype=torch.uint8, device=x.device)
    terms = filt.terms if filt is not None else 3
    stats = filt.stats if filt is not None else None
    check(lib.qk_kmeans_assign_filtered(ptr(x), n, x.stride(0), d, ptr(centroids), K, centroids.stride(0), metric, terms,
                                        ptr(out), ptr(dist), ptr(stats), ptr(ws), wsb, _stream()))
    if filt is not None:
        filt.points += n
    return (out, dist) if want_dist else out


def partition_by_assignment(assign: torch.Tensor, K: int):
    """counts [K], offsets [K+1], order [n] (int64, device): rows grouped by centroid, ascending row
    index inside a group."""
    lib = _lib.load()
    n = int(assign.shape[0])
    dev = assign.device
    counts = torch.zeros(K, dtype=torch.int64, device=dev)
    offsets = torch.zeros(K + 1, dtype=torch.int64, device=dev)
    order = torch.empty(n, dtype=torch.int64, device=dev)
    wsb = lib.qk_partition_workspace_bytes(n, K)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    check(lib.qk_partition_by_assignment(ptr(assign), n, K, ptr(counts), ptr(offsets), ptr(order), ptr(ws), wsb,
                                         _stream()))
    return counts, offsets, order


def centroid_sums(x: torch.Tensor, d: int, order: torch.Tensor, offsets: torch.Tensor, K: int) -> torch.Tensor:
    lib = _lib.load()
    pitch = (d + 3) // 4 * 4
    sums = torch.zeros((K, pitch), dtype=torch.float32, device=x.device)
    check(lib.qk_kmeans_accumulate(ptr(x), x.stride(0), d, ptr(order), ptr(offsets), K, ptr(sums), pitch, _stream()))
    return sums


def train_centroids(x: torch.Tensor, d: int, K: int, metric: int, niter: int, filt: AssignFilter | None = None) -> torch.Tensor:
    """faiss::Clustering::train (Clustering.cpp:255-539) with default ClusteringParameters:
    subsample to K*256 points (seed 1234), initial centroids = first K of rand_perm(seed + 1),
    niter x [assign; mean update; split empty clusters]."""
    lib = _lib.load()
    n = int(x.shape[0])
    if n < K:
        raise RuntimeError(f"Number of training points ({n}) should be at least as large as number of clusters ({K})")
    xt = x
    if n > K * MAX_POINTS_PER_CENTROID:
        perm = rand_perm_prefix(n, FAISS_SEED, K * MAX_POINTS_PER_CENTROID)
        xt = x[torch.from_numpy(perm).to(x.device)]
    nx = int(xt.shape[0])
    if nx == K:
        return x[:K].clone()
    perm = rand_perm_prefix(nx, FAISS_SEED + 1, K)
    centroids = xt[torch.from_numpy(perm).to(x.device)].clone()
    filt = filt if filt is not None else AssignFilter(x.device)
    for _ in range(niter):
        assign = assign_points(xt, d, centroids, metric, filt=filt)
        counts, offsets, order = partition_by_assignment(assign, K)
        sums = centroid_sums(xt, d, order, offsets, K)
        hassign = counts.to(torch.float32)
        inv = torch.where(hassign > 0, 1.0 / hassign, torch.zeros_like(hassign))
        centroids = sums * inv[:, None]
        filt.review()  # beside the synchronisation below
        if bool((counts == 0).any()):
            c_h = centroids.cpu().contiguous()
            h_h = hassign.cpu().contiguous()
            nsplit = C.c_int64(0)
            check(lib.qk_host_split_clusters(d, K, nx, h_h.numpy().ctypes.data_as(C.POINTER(C.c_float)),
                                             c_h.numpy().ctypes.data_as(C.POINTER(C.c_float)), c_h.stride(0),
                                             C.byref(nsplit)))
            centroids = c_h.to(x.device)
    return centroids


def kmeans(x: torch.Tensor, d: int, K: int, metric: int, niter: int):
    """clustering.cpp:13-97. `x` is our own [n, pitch] device copy and is normalised IN PLACE for the
    inner-product metric (the reference stores the normalised vectors, clustering.cpp:25-26).
    Returns (centroids [K, pitch], counts, offsets, order)."""
    lib = _lib.load()
    n = int(x.shape[0])
    if metric == _lib.QK_METRIC_INNER_PRODUCT:
        check(lib.qk_normalize_rows(ptr(x), n, x.stride(0), d, _stream()))
    filt = AssignFilter(x.device)
    trained = train_centroids(x, d, K, metric, niter, filt)
    centroids = trained
    if metric == _lib.QK_METRIC_INNER_PRODUCT:
        centroids = trained.clone()
        check(lib.qk_normalize_rows(ptr(centroids), K, centroids.stride(0), d, _stream()))
    # the final assignment searches the faiss index, which still holds the centroids of the last
    # iteration -- un-normalised for ip (clustering.cpp:65 uses index_ptr, not the normalised copy
    # that is returned); argmax <x, c> and argmax <x, c/|c|> can differ, so keep the reference's choice.
    assign = assign_points(x, d, trained, metric, filt=filt)
    counts, offsets, order = partition_by_assignment(assign, K)
    return centroids, counts, offsets, order


def kmeans_refine(centroids: torch.Tensor, d: int, vecs: torch.Tensor, ids: torch.Tensor, metric: int,
                  iterations: int):
    """kmeans_refine_partitions (clustering.cpp:99-182) over the members of the selected partitions.

    vecs [n, pitch] / ids [n] hold the members concatenated partition by partition -- the order in which
    the reference visits them (clustering.cpp:141-176). Every iteration re-concatenates the members
    cluster by cluster in visiting order, exactly like the reference's rebuilt partitions, so both the
    per-cluster accumulation order (clustering.cpp:168-173) and the final list contents match.
    Returns (centroids used for the LAST assignment, counts [K] int64, vecs, ids) with the members of
    cluster c at rows [cumsum(counts)[c-1], cumsum(counts)[c])."""
    K = int(centroids.shape[0])
    iters = iterations if iterations > 0 else 1
    n = int(vecs.shape[0])
    counts = torch.zeros(K, dtype=torch.int64, device=vecs.device)
    identity = torch.arange(n, dtype=torch.int64, device=vecs.device)
    offsets = None
    filt = AssignFilter(vecs.device)
    for it in range(iters):
        filt.review()
        if it > 0:
            sums = centroid_sums(vecs, d, identity, offsets, K)
            # 0/0 = NaN for an emptied cluster, as in the reference (clustering.cpp:122-124)
            centroids = sums / counts.to(torch.float32)[:, None]
        if n == 0:
            offsets = torch.zeros(K + 1, dtype=torch.int64, device=vecs.device)
            continue
        assign = assign_points(vecs, d, centroids, metric, filt=filt)
        counts, offsets, order = partition_by_assignment(assign, K)
        vecs = vecs[order]
        ids = ids[order]
    return centroids, counts, vecs, ids
