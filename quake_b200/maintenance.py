"""Maintenance policy: hit rates -> cost model -> delete / split decisions -> local refinement.

Host-side mirror of the reference's ``MaintenancePolicy`` + ``MaintenanceCostEstimator`` + ``HitCountTracker``
(/root/reference/src/cpp/src/maintenance_policies.cpp:33-202, src/maintenance_cost_estimator.cpp:131-493,
src/hit_count_tracker.cpp:43-66), with the same decision rules and parameter names. Two things differ by design:

* the latency function lambda(n, k) the cost model is built on is MEASURED ON THIS GPU (the reference profiles its CPU
  ``scan_list``, maintenance_cost_estimator.cpp:59-94): the amortised time one query of a batch spends scanning a
  list of n vectors with a top-k of k, on the same grid of (n, k) points and with the same bilinear inter- /
  extrapolation (:131-251);
* hits are recorded by ``QuakeIndex.search`` itself (the reference's coordinator never calls ``record_query_hits`` in
  this snapshot, so its ``maintenance()`` has nothing to act on from Python); the window keeps the last
  ``window_size`` queries like ``HitCountTracker``.

The heavy lifting (k-means with K = 2 per split partition, the refit of the neighbourhood, re-assignment of deleted
partitions' vectors) runs on the device through the same kernels as build(): clustering.kmeans / kmeans_refine and
the coarse scan.
"""
from __future__ import annotations

import time

import numpy as np
import torch

from .params import MaintenancePolicyParams, MaintenanceTimingInfo, SearchParams

LATENCY_RANGE_N = [1, 2, 4, 16, 64, 256, 1024, 4096, 16384, 65536]  # common.h:97
LATENCY_RANGE_K = [1, 4, 16, 64, 256]                                # common.h:98
LATENCY_NTRIALS = 5                                                  # common.h:99
_PROFILE_QUERIES = 256

_latency_models: dict = {}


class ListScanLatencyEstimator:
    """lambda(n, k) in nanoseconds on a grid, bilinear interpolation inside it and linear extrapolation beyond
    (maintenance_cost_estimator.cpp:131-251)."""

    def __init__(self, d: int, n_values=None, k_values=None, table: np.ndarray | None = None, device=None):
        self.d = int(d)
        self.n_values = list(n_values or LATENCY_RANGE_N)
        self.k_values = list(k_values or LATENCY_RANGE_K)
        if sorted(self.n_values) != self.n_values:
            raise RuntimeError("n_values must be sorted in ascending order.")
        if sorted(self.k_values) != self.k_values:
            raise RuntimeError("k_values must be sorted in ascending order.")
        self.model = np.asarray(table, dtype=np.float32) if table is not None else self._profile(device)

    def _profile(self, device) -> np.ndarray:
        """Measure the partition scan on this GPU: a flat store of n vectors, a batch of 256 queries, top-k; lambda =
        batch time / 256 (what one more (query, list) pair of a batched search costs)."""
        from .index import QuakeIndex, scan_partitions
        from .params import IndexBuildParams
        from . import clustering
        g = torch.Generator().manual_seed(77)
        max_n = self.n_values[-1]
        vectors = torch.rand(max_n, self.d, generator=g)
        q = clustering.pad_rows(torch.rand(_PROFILE_QUERIES, self.d, generator=g), device)
        out = np.zeros((len(self.n_values), len(self.k_values)), dtype=np.float32)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i, n in enumerate(self.n_values):
            idx = QuakeIndex()
            idx.build(vectors[:n], torch.arange(n, dtype=torch.int64), IndexBuildParams())
            for j, k in enumerate(self.k_values):
                scan_partitions(idx.store, q, None, k, idx.metric)  # warm-up
                e0.record()
                for _ in range(LATENCY_NTRIALS):
                    scan_partitions(idx.store, q, None, k, idx.metric)
                e1.record()
                torch.cuda.synchronize()
                out[i, j] = e0.elapsed_time(e1) * 1e6 / LATENCY_NTRIALS / _PROFILE_QUERIES
        return out

    @staticmethod
    def _axis(values, target):
        """(lower index, upper index, fraction as float32). Inside the grid the fraction is the position between the two
        grid lines; beyond it, the reference's extrapolation fraction (target - last) / (last - second last)."""
        f32 = np.float32
        if target <= values[-1]:
            upper = int(np.searchsorted(values, target, side="right"))
            if upper >= len(values):
                return len(values) - 2, len(values) - 1, f32(1.0)
            lower = upper - 1
            return lower, upper, f32(target - values[lower]) / f32(values[upper] - values[lower])
        lower, upper = len(values) - 2, len(values) - 1
        return lower, upper, f32(target - values[upper]) / f32(values[upper] - values[lower])

    def estimate_scan_latency(self, n: int, k: int) -> np.float32:
        """Single-precision arithmetic in the reference's operation order (:131-251): the policy compares the deltas
        built from these values with thresholds, so the last bits decide borderline partitions."""
        n, k = int(n), int(k)
        f32 = np.float32
        if n == 0 or k == 0:
            return f32(0.0)
        if n < self.n_values[0] or k < self.k_values[0]:
            raise IndexError("n or k is below the minimum supported values.")
        n_within, k_within = n <= self.n_values[-1], k <= self.k_values[-1]
        il, iu, t = self._axis(self.n_values, n)
        jl, ju, u = self._axis(self.k_values, k)
        m = self.model
        f11, f12, f21, f22 = f32(m[il, jl]), f32(m[il, ju]), f32(m[iu, jl]), f32(m[iu, ju])
        one = f32(1.0)

        def extrap(f1, f2, frac):  # linear_extrapolate (maintenance_cost_estimator.h:123-126)
            return f2 + (f2 - f1) * frac

        if n_within and k_within:
            return (one - t) * (one - u) * f11 + t * (one - u) * f21 + (one - t) * u * f12 + t * u * f22
        if not n_within and k_within:
            return (one - u) * extrap(f11, f21, t) + u * extrap(f12, f22, t)
        if n_within and not k_within:
            return (one - t) * extrap(f11, f12, u) + t * extrap(f21, f22, u)
        return extrap(extrap(f11, f21, t), extrap(f12, f22, t), u)


def latency_model(d: int, device) -> ListScanLatencyEstimator:
    key = (int(d), str(device))
    if key not in _latency_models:
        _latency_models[key] = ListScanLatencyEstimator(d, device=device)
    return _latency_models[key]


class MaintenanceCostEstimator:
    """maintenance_cost_estimator.cpp:355-493 -- the three deltas, verbatim in structure; k is fixed at 10 like the
    reference's MaintenancePolicy constructor (maintenance_policies.cpp:24-27)."""

    def __init__(self, d: int, alpha: float, k: int, latency: ListScanLatencyEstimator):
        if d <= 0:
            raise ValueError("Dimension must be positive")
        if k <= 0:
            raise ValueError("k must be positive")
        if alpha <= 0.0:
            raise ValueError("alpha must be positive")
        self.d, self.alpha, self.k, self.latency = d, float(alpha), int(k), latency

    def L(self, n) -> np.float32:
        return self.latency.estimate_scan_latency(int(n), self.k)

    # the three deltas in float32 and in the reference's operation order (borderline decisions depend on the last bits)
    def compute_split_delta(self, partition_size: int, hit_rate: float, total_partitions: int) -> float:
        f32 = np.float32
        hit_rate = f32(hit_rate)
        delta_overhead = self.L(total_partitions + 1) - self.L(total_partitions)
        old_cost = self.L(partition_size) * hit_rate
        new_cost = self.L(partition_size // 2) * hit_rate * (f32(2.0) * f32(self.alpha))
        return float(delta_overhead + new_cost - old_cost)

    def compute_delete_delta(self, partition_size: int, hit_rate: float, total_partitions: int,
                             avg_partition_hit_rate: float, avg_partition_size: float) -> float:
        if total_partitions <= 1:
            return 0.0
        f32 = np.float32
        T = int(total_partitions)
        hit_rate, avg_rate, avg_size = f32(hit_rate), f32(avg_partition_hit_rate), f32(avg_partition_size)
        delta_overhead = self.L(T - 1) - self.L(T)
        cost_old = f32(T - 1) * avg_rate * self.L(int(avg_size)) + hit_rate * self.L(partition_size)
        merged_size = avg_size + f32(partition_size) / f32(T - 1)
        merged_hit_rate = avg_rate + hit_rate / f32(T - 1)
        if partition_size < T:
            cost_new = (f32(partition_size) * merged_hit_rate * self.L(int(avg_size + f32(1.0)))
                        + f32(T - partition_size - 1) * merged_hit_rate * self.L(int(avg_size)))
        else:
            cost_new = f32(T - 1) * merged_hit_rate * self.L(int(np.ceil(merged_size)))
        return float(delta_overhead + (cost_new - cost_old))

    def compute_delete_delta_w_reassign(self, partition_size: int, hit_rate: float, total_partitions: int,
                                        reassign_counts, reassign_sizes, reassign_hit_rates) -> float:
        if total_partitions <= 1:
            return 0.0
        f32 = np.float32
        T = int(total_partitions)
        hit_rate = f32(hit_rate)
        delta_overhead = self.L(T - 1) - self.L(T)
        removal_delta = hit_rate * self.L(partition_size)
        reassign_delta = f32(0.0)
        for _cnt, size, rate in zip(reassign_counts, reassign_sizes, reassign_hit_rates):
            rate = f32(rate)
            old = rate * self.L(size)
            new_size = f32(size + partition_size)
            reassign_delta = reassign_delta + ((rate + hit_rate) * self.L(int(new_size)) - old)
        return float(delta_overhead + removal_delta + reassign_delta)


class HitCountTracker:
    """hit_count_tracker.cpp:43-66 for batches: the last `window_size` queries' probed partitions (device tensors,
    one [Q, nprobe] block per search) and their scanned fractions."""

    def __init__(self, window_size: int, total_vectors: int):
        if window_size <= 0:
            raise ValueError("Window size must be positive")
        if total_vectors <= 0:
            raise ValueError("Total vectors must be positive")
        self.window_size = int(window_size)
        self.total_vectors = int(total_vectors)
        self.reset()

    def reset(self) -> None:
        self.blocks: list = []  # (partition ids [q, p] (-1 = none), partitions scanned [q] or None) per search
        self.num_queries_recorded = 0

    def add_batch(self, p_ids: torch.Tensor, scanned: torch.Tensor | None) -> None:
        """p_ids [Q, P] probed partition ids in rank order; scanned [Q] = how many of them each query actually scanned
        (APS) or None = all. O(1) on the search path: the block is kept as is (a device slice), everything else is
        computed when maintenance() asks for the window."""
        w = self.window_size
        self.blocks.append((p_ids[-w:].detach(), None if scanned is None else scanned[-w:].detach()))
        total = sum(int(b[0].shape[0]) for b in self.blocks)
        while len(self.blocks) > 1 and total - int(self.blocks[0][0].shape[0]) >= w:
            total -= int(self.blocks.pop(0)[0].shape[0])
        self.num_queries_recorded = min(total, w)

    def window(self, sizes_by_pid: torch.Tensor):
        """(hit partition ids of the window's queries [n, P] with -1 padding; number of queries; mean scanned
        fraction = HitCountTracker::get_current_scan_fraction, with the partitions' CURRENT sizes -- the reference
        snapshots them per query)."""
        w = self.window_size
        if not self.blocks:
            return torch.zeros((0, 1), dtype=torch.int64), 0, 1.0
        P = max(int(b[0].shape[1]) for b in self.blocks)
        parts = []
        for ids, scanned in self.blocks:
            if scanned is not None:  # APS: only the first `scanned` candidates of a query were scanned
                rank = torch.arange(ids.shape[1], device=ids.device)[None, :]
                ids = torch.where(rank < scanned[:, None].to(torch.int64), ids, torch.full_like(ids, -1))
            if int(ids.shape[1]) < P:  # searches with different nprobe in the window: pad to the widest
                pad = torch.full((ids.shape[0], P - ids.shape[1]), -1, dtype=ids.dtype, device=ids.device)
                ids = torch.cat([ids, pad], 1)
            parts.append(ids)
        ids = torch.cat(parts, 0)[-w:]
        valid = (ids >= 0) & (ids < sizes_by_pid.numel())
        sizes = torch.where(valid, sizes_by_pid[ids.clamp(0, max(int(sizes_by_pid.numel()) - 1, 0))], torch.zeros_like(ids))
        frac = sizes.to(torch.float32).sum(1) / float(self.total_vectors)
        return ids, int(ids.shape[0]), float(frac.mean().item()) if frac.numel() else 1.0


class MaintenancePolicy:
    def __init__(self, index, params: MaintenancePolicyParams):
        self.index = index
        self.params = params
        self.tracker = HitCountTracker(int(params.window_size), max(index.ntotal(), 1))
        self._estimator = None

    @property
    def cost_estimator(self) -> MaintenanceCostEstimator:
        if self._estimator is None:  # the latency profile is measured on first use (a few ms), once per (d, device)
            st = self.index.store
            self._estimator = MaintenanceCostEstimator(st.d, float(self.params.alpha), 10, latency_model(st.d, st.device))
        return self._estimator

    def record_query_hits(self, p_ids: torch.Tensor, scanned: torch.Tensor | None = None) -> None:
        self.tracker.add_batch(p_ids, scanned)

    def reset(self) -> None:
        self.tracker.reset()

    # ------------------------------------------------------------------ maintenance_policies.cpp:33-172
    def perform_maintenance(self) -> MaintenanceTimingInfo:
        p = self.params
        idx = self.index
        info = MaintenanceTimingInfo()
        recorded = self.tracker.num_queries_recorded
        if idx.parent is None or recorded < int(p.window_size):
            print(f"Window not full yet. {recorded} queries recorded and {p.window_size} queries required.")
            return info
        t_total = time.perf_counter()
        st = idx.store
        # STEP 1: aggregate hit counts
        hit_ids, _nq, scan_fraction = self.tracker.window(st.device_sizes_by_pid())
        flat = hit_ids.reshape(-1)
        flat = flat[flat >= 0]
        nbins = int(st.curr_list_id)
        hits = torch.bincount(flat, minlength=nbins).cpu().numpy() if flat.numel() else np.zeros(nbins, np.int64)
        all_pids = st.partition_ids()
        # STEP 2: decisions
        T = idx.nlist()
        avg_size = idx.ntotal() // max(T, 1)
        est = self.cost_estimator
        to_delete, to_split = [], []
        for pid in all_pids.tolist():
            hit_rate = float(hits[pid]) / float(p.window_size) if pid < hits.size else 0.0
            size = st.size_of(pid)
            delete_delta = est.compute_delete_delta(size, hit_rate, T, scan_fraction, avg_size)
            if delete_delta < -float(p.delete_threshold_ns):
                if p.enable_delete_rejection and size > int(p.min_partition_size):
                    if self._delete_with_reassign_delta(pid, size, hit_rate, T, hits) < -float(p.delete_threshold_ns):
                        to_delete.append(pid)
                else:
                    to_delete.append(pid)
            elif size > int(p.min_partition_size):
                if est.compute_split_delta(size, hit_rate, T) < -float(p.split_threshold_ns):
                    to_split.append(pid)
        # STEP 3: deletions (vectors re-assigned to their nearest remaining partition)
        t0 = time.perf_counter()
        if to_delete:
            idx.delete_partitions(torch.tensor(to_delete, dtype=torch.int64), reassign=True)
        torch.cuda.synchronize()
        info.delete_time_us = int((time.perf_counter() - t0) * 1e6)
        # STEP 4: splits
        t0 = time.perf_counter()
        new_pids = None
        if to_split:
            ts = torch.tensor(to_split, dtype=torch.int64)
            split = idx.split_partitions(ts)
            idx.delete_partitions(ts, reassign=False)
            new_pids = idx.add_partitions(split)
        torch.cuda.synchronize()
        info.split_time_us = int((time.perf_counter() - t0) * 1e6)
        # STEP 5: local refinement around the new partitions
        t0 = time.perf_counter()
        if new_pids is not None and new_pids.numel() > 0:
            self.local_refinement(new_pids)
        torch.cuda.synchronize()
        info.split_refine_time_us = int((time.perf_counter() - t0) * 1e6)
        info.n_splits, info.n_deletes = len(to_split), len(to_delete)
        info.total_time_us = int((time.perf_counter() - t_total) * 1e6)
        return info

    def _delete_with_reassign_delta(self, pid: int, size: int, hit_rate: float, T: int, hits: np.ndarray) -> float:
        """maintenance_policies.cpp:83-125: where would this partition's vectors go (second nearest centroid -- the
        nearest is the partition itself), and what does that do to the scan cost?"""
        idx = self.index
        sp = SearchParams()
        sp.k, sp.batched_scan = 2, True
        vecs, _ = idx.store.get_list(pid, padded=True)
        ids, _, _ = idx.parent._search_device(vecs.contiguous(), sp)
        r = ids.reshape(-1)
        r = r[(r != pid) & (r >= 0)]
        uniq, counts = torch.unique(r, return_counts=True)
        uniq_h, counts_h = uniq.cpu().tolist(), counts.cpu().tolist()
        sizes = [idx.store.size_of(u) for u in uniq_h]
        rates = [float(hits[u]) / float(self.params.window_size) if u < hits.size else 0.0 for u in uniq_h]
        return self.cost_estimator.compute_delete_delta_w_reassign(size, hit_rate, T, counts_h, sizes, rates)

    def local_refinement(self, partition_ids: torch.Tensor) -> None:
        """maintenance_policies.cpp:183-202: refit the `refinement_radius` nearest partitions of every new centroid."""
        p = self.params
        if int(p.refinement_radius) == 0:
            return
        idx = self.index
        cents = idx.parent.get(partition_ids)
        sp = SearchParams()
        sp.nprobe, sp.k = 1000, int(p.refinement_radius)
        res = idx.parent.search(cents, sp)
        ids = torch.unique(res.ids)
        ids = ids[ids != -1]
        idx.refine_partitions(ids, int(p.refinement_iterations))
