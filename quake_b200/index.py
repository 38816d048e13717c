"""QuakeIndex: the reference's Python-facing index object, hosted on one B200.

Mirrors ``QuakeIndex`` + ``QueryCoordinator`` + ``PartitionManager`` of the reference for the hot path
(/root/reference/src/cpp/src/quake_index.cpp, query_coordinator.cpp:612-657, partition_manager.cpp) with
the method names, argument meaning, output conventions and error classes of the pybind11 surface
(/root/reference/src/cpp/bindings/wrap.cpp:57-128). Distance arithmetic and top-k run in the CUDA
kernels behind include/quake_b200.h; nothing here falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C
import os
import struct
import time

import numpy as np
import torch

from . import _lib, aps, clustering
from ._lib import check, ptr
from .params import (BuildTimingInfo, IndexBuildParams, MaintenancePolicyParams, MaintenanceTimingInfo,
                     ModifyTimingInfo, SearchParams, SearchResult, SearchTimingInfo, str_to_metric)
from .store import PartitionStore

SERIALIZATION_MAGIC = 0x44494E4C  # common.h:66
SERIALIZATION_VERSION = 3         # common.h:67
_WORKSPACE_LIMIT = 6 << 30        # split query batches whose scan workspace would exceed this


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _device() -> torch.device:
    _lib.require_device()
    return torch.device("cuda", torch.cuda.current_device())


# Filter precision policy of the PARTITION scan of a two-level index ("auto" | "2" | "3", env QK_FILTER): the
# tensor-core filter may drop the a_lo term (2xTF32: a few % faster scan kernel). Results stay exact either way -- the
# refine step proves every answer and re-scans a query exactly when the looser filter cannot separate its k-th
# neighbour -- but re-scans are slow (one CTA walks all probed rows), so "auto" starts SAFE at 3 terms; every 3-term
# scan also counts the queries whose proof would have failed under the 2-term bound, and once >= 256 queries have been
# seen with at most 0.5 % such failures the store switches to 2 terms. A 2-term search that reports more than 1 % of
# its queries re-scanned switches back for good. All of it from the asynchronously read statistics of earlier calls.
# Coarse scans, flat indexes, APS and the k-means assign always run 3xTF32 (their top-k margins are much tighter).
FILTER_POLICY = os.environ.get("QK_FILTER", "auto")
LAST_SCAN_STATS = None  # set QK_SCAN_STATS=1: int32[4] device tensor of the last qk_scan_partitions call

# Fixed-nprobe searches are replayed from a CUDA graph captured per (batch size, k, nprobe, index version): the
# step is ~20 short launches, and the gaps between them cost as much as a kernel. QK_GRAPH=0 (or GRAPHS_ENABLED =
# False) keeps every call eager -- bench.py does that while it times individual kernels with events.
GRAPHS_ENABLED = os.environ.get("QK_GRAPH", "1") != "0"
_MAX_PLANS = 8
_GRAPH_MAX_Q = 16384  # larger batches run eagerly: their launches are long, and a plan pins its whole workspace
_capturing = False
GRAPH_LAUNCHES = 0  # kernels of ours launched through graph replays (qk_launch_count() only sees host-side launches)


def launch_count() -> int:
    """Kernels of this library launched so far: host-side launches + those inside replayed CUDA graphs."""
    return int(_lib.load().qk_launch_count()) + GRAPH_LAUNCHES


class _FilterMonitor:
    """Asynchronous read-back of a search's scan statistics (32 bytes through pinned memory + an event; never waited
    for): how many queries of an earlier batch fell into the exact re-scan / would have under the 2-term filter."""

    def __init__(self):
        self.host = torch.zeros(8, dtype=torch.int32).pin_memory()
        self.event = None
        self.queries = 0
        self.calls = 0
        self.seen = 0         # queries observed in 3-term mode ...
        self.would_fail = 0   # ... and how many of them the 2-term bound would have re-scanned
        self.locked = False   # a 2-term search re-scanned too much: 3 terms for good

    def submit(self, stats: torch.Tensor, queries: int) -> None:
        self.calls += 1
        if self.event is not None or stats is None or self.locked:
            return  # one read-back in flight at a time
        if self.calls > 32 and (self.calls & 7):
            return  # steady state: every 8th call
        self.host.copy_(stats, non_blocking=True)
        self.event = torch.cuda.Event()
        self.event.record()
        self.queries = queries

    def poll(self):
        """(queries re-scanned, would-fail-under-2-terms, queries) of a finished read-back, or None."""
        if self.event is None or not self.event.query():
            return None
        self.event = None
        return int(self.host[0]), int(self.host[4]), self.queries


class _SearchPlan:
    """One captured search: static input buffer -> static output buffers."""

    def __init__(self, index: "QuakeIndex", Q: int, sp: SearchParams):
        global _capturing
        dev = index.store.device
        self.xq = torch.zeros((Q, index.store.pitch), dtype=torch.float32, device=dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        _capturing = True
        try:
            with torch.cuda.stream(side):
                for _ in range(2):  # warm-up outside the capture: lazy tables, function attributes, allocator pools
                    index._search_core(self.xq, sp)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            lib = _lib.load()
            n0 = lib.qk_launch_count()
            # thread_local: other threads (e.g. the NCCL watchdog of a multi-rank job) may touch the CUDA runtime
            with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
                self.ids, self.dist, self.p_ids = index._search_core(self.xq, sp)
            self.launches = int(lib.qk_launch_count() - n0)  # kernels of ours one replay launches
            self.stats = index.__dict__.get("_last_stats")   # written by the replay (static buffer of the plan)
        finally:
            _capturing = False
        # everything outside the graph's private pool whose address the graph baked in lives as long as the plan: the
        # flat probe tables (this index's or its parent's) and the stores' segment tables / arenas of this version
        self.keep = [index._flat_probe(Q) if index.parent is None else None,
                     index.parent._flat_probe(Q) if index.parent is not None and index.parent.parent is None else None,
                     index.store.tables_snapshot(),
                     index.parent.store.tables_snapshot() if index.parent is not None else None]

    def run(self, xq: torch.Tensor):
        if xq.data_ptr() != self.xq.data_ptr():
            self.xq.copy_(xq, non_blocking=True)
        self.graph.replay()
        global GRAPH_LAUNCHES
        GRAPH_LAUNCHES += self.launches
        return self.ids, self.dist, self.p_ids


def scan_partitions(store: PartitionStore, xq: torch.Tensor, probe_slots: torch.Tensor | None, k: int, metric: int,
                    want_rows: bool = False, filter_terms: int | None = None):
    """qk_scan_partitions over a device query batch xq [Q, pitch] and probe_slots [Q, nprobe] (int32), or None for a
    single-list store (flat mode: every query scans the whole list, no probe table, no grouping kernels).
    Returns (ids [Q,k] int64, distances [Q,k] float32[, rows]) on the device."""
    lib = _lib.load()
    if probe_slots is None:
        if store.slot_pid.size != 1:
            raise RuntimeError("quake_b200: flat-mode scan needs a single-list store")
        Q, nprobe = int(xq.shape[0]), 1
    else:
        Q, nprobe = int(probe_slots.shape[0]), int(probe_slots.shape[1])
    st, _ = store.tables(store.segment_len(Q, nprobe))
    if filter_terms is not None and int(filter_terms) != st.filter_terms:
        keep = st
        st = _lib.QkStore.from_buffer_copy(st)  # same device tables, other filter precision (callers without a monitor)
        st.filter_terms = int(filter_terms)
        st._keepalive = keep
    dev = xq.device
    out_ids = torch.empty((Q, k), dtype=torch.int64, device=dev)
    out_dist = torch.empty((Q, k), dtype=torch.float32, device=dev)
    out_rows = torch.empty((Q, k), dtype=torch.int64, device=dev) if want_rows else None
    if st.num_segments == 0:
        out_ids.fill_(-1)
        out_dist.fill_(float("-inf") if metric == _lib.QK_METRIC_INNER_PRODUCT else float("inf"))
        if want_rows:
            out_rows.fill_(-1)
        return (out_ids, out_dist, out_rows) if want_rows else (out_ids, out_dist)
    if probe_slots is not None:
        probe_slots = probe_slots.to(torch.int32).contiguous()
    # chunk the batch so that the workspace stays bounded
    chunk = Q
    while True:
        wsb = lib.qk_scan_workspace_bytes(C.byref(st), chunk, nprobe, k)
        if wsb == 0:
            check(lib.qk_scan_partitions(C.byref(st), ptr(xq), chunk, xq.stride(0), ptr(probe_slots), nprobe, metric, k,
                                         ptr(out_ids), ptr(out_dist), None, None, 0, None, _stream()))
            raise _lib.QuakeB200Error("scan plan failed")
        if wsb <= _WORKSPACE_LIMIT or chunk <= 32:
            break
        chunk = (chunk + 1) // 2
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    stats = None
    if os.environ.get("QK_SCAN_STATS") == "1":
        global LAST_SCAN_STATS
        stats = LAST_SCAN_STATS = torch.zeros(8, dtype=torch.int32, device=dev)
    for b in range(0, Q, chunk):
        n = min(chunk, Q - b)
        check(lib.qk_scan_partitions(C.byref(st), ptr(xq[b:]), n, xq.stride(0),
                                     ptr(probe_slots[b:]) if probe_slots is not None else None, nprobe, metric, k,
                                     ptr(out_ids[b:]), ptr(out_dist[b:]), ptr(out_rows[b:]) if want_rows else None,
                                     ptr(ws), wsb, ptr(stats), _stream()))
    return (out_ids, out_dist, out_rows) if want_rows else (out_ids, out_dist)


class QuakeIndex:
    def __init__(self, current_level: int = 0):
        self.parent: "QuakeIndex | None" = None
        self.current_level = int(current_level)
        self.store: PartitionStore | None = None
        self.metric = 1
        self.build_params: IndexBuildParams | None = None
        self.maintenance_policy_params: MaintenancePolicyParams | None = None
        self.maintenance_policy = None  # maintenance.MaintenancePolicy (hit window + cost model), level 0 only

    # ------------------------------------------------------------------ build (quake_index.cpp:29-90)
    def build(self, x: torch.Tensor, ids: torch.Tensor, build_params: IndexBuildParams) -> BuildTimingInfo:
        t0 = time.perf_counter()
        dev = _device()
        self.build_params = build_params
        self.metric = str_to_metric(build_params.metric)
        if x.dim() != 2:
            raise RuntimeError("[QuakeIndex::build] x must be 2D [N, dim]")
        n, d = int(x.shape[0]), int(x.shape[1])
        if ids.shape[0] != n:
            raise RuntimeError("[QuakeIndex::build] x and ids must have the same number of rows")
        info = BuildTimingInfo()
        info.n_vectors, info.d = n, d
        xd = clustering.pad_rows(x, dev)
        if xd.data_ptr() == x.data_ptr():
            xd = xd.clone()  # the reference deep-copies x (quake_index.cpp:33)
        idd = ids.to(device=dev, dtype=torch.int64).contiguous()
        self.store = PartitionStore(d, dev)
        self._reset_caches()
        nlist = int(build_params.nlist)
        if nlist > 1:
            s1 = time.perf_counter()
            centroids, counts, offsets, order = clustering.kmeans(xd, d, nlist, self.metric, int(build_params.niter))
            torch.cuda.synchronize()
            info.train_time_us = int((time.perf_counter() - s1) * 1e6)
            s2 = time.perf_counter()
            self.parent = QuakeIndex(self.current_level + 1)
            pp = IndexBuildParams()
            pp.metric = build_params.metric
            pp.num_workers = build_params.num_workers
            self.parent.build(centroids[:, :d], torch.arange(nlist, dtype=torch.int64), pp)
            self.store.init_from_sorted(xd, idd, order, counts.cpu().numpy(), np.arange(nlist, dtype=np.int64))
            torch.cuda.synchronize()
            info.assign_time_us = int((time.perf_counter() - s2) * 1e6)
        else:
            # flat index: one partition with id 0 (quake_index.cpp:66-79)
            self.parent = None
            self.store.init_from_sorted(xd, idd, None, np.array([n], dtype=np.int64), np.array([0], dtype=np.int64))
        info.n_clusters = self.nlist()
        self.initialize_maintenance_policy(MaintenancePolicyParams())
        self._apply_filter_policy()
        torch.cuda.synchronize()
        info.total_time_us = int((time.perf_counter() - t0) * 1e6)
        return info

    @classmethod
    def from_partitions(cls, centroids: torch.Tensor, vectors: list, ids: list, metric: str = "l2") -> "QuakeIndex":
        """A two-level index over GIVEN partitions: partition j holds vectors[j] / ids[j] and is represented by
        centroids[j] (PartitionManager::init_partitions with a ready-made Clustering, partition_manager.cpp:33-121)."""
        dev = _device()
        self = cls()
        self.metric = str_to_metric(metric)
        self.build_params = IndexBuildParams()
        self.build_params.metric = metric
        self.build_params.nlist = len(vectors)
        d = int(centroids.shape[1])
        counts = np.array([int(v.shape[0]) for v in vectors], dtype=np.int64)
        xd = clustering.pad_rows(torch.cat([v.reshape(-1, d) for v in vectors]), dev)
        idd = torch.cat([i.reshape(-1) for i in ids]).to(device=dev, dtype=torch.int64)
        self.store = PartitionStore(d, dev)
        self.store.init_from_sorted(xd, idd, None, counts, np.arange(len(vectors), dtype=np.int64))
        self.parent = cls(1)
        pp = IndexBuildParams()
        pp.metric = metric
        self.parent.build(centroids, torch.arange(len(vectors), dtype=torch.int64), pp)
        self.initialize_maintenance_policy(MaintenancePolicyParams())
        self._apply_filter_policy()
        return self

    # ------------------------------------------------------------------ search (query_coordinator.cpp:612-657)
    def _check_built(self, who: str):
        if self.store is None:
            raise RuntimeError(f"[QuakeIndex::{who}()] No query coordinator. Did you build the index?")

    def search(self, x: torch.Tensor, search_params: SearchParams) -> SearchResult:
        self._check_built("search")
        t0 = time.perf_counter()
        res = SearchResult()
        tinfo = SearchTimingInfo()
        tinfo.search_params = search_params
        tinfo.n_clusters = self.nlist()
        res.timing_info = tinfo
        if x is None or x.numel() == 0 or x.shape[0] == 0:
            # query_coordinator.cpp:476-482
            res.ids = torch.empty((0,), dtype=torch.int64)
            res.distances = torch.empty((0,), dtype=torch.float32)
            tinfo.parent_info = SearchTimingInfo()
            return res
        if x.dim() != 2 or int(x.shape[1]) != self.store.d:
            raise RuntimeError(f"[QuakeIndex::search] queries must be [Q, {self.store.d}]")
        out_dev = x.device
        use_aps = self.parent is not None and float(search_params.recall_target) > 0.0 and not bool(search_params.batched_scan)
        if (GRAPHS_ENABLED and not use_aps and self.current_level == 0 and x.dtype == torch.float32
                and int(x.shape[0]) <= _GRAPH_MAX_Q):
            # straight into the plan's static input buffer (one H2D copy when x is a host tensor)
            plan = self._plan(int(x.shape[0]), search_params)
            if plan is not None:
                xq = plan.xq
                xq[:, : self.store.d].copy_(x, non_blocking=True)
            else:
                xq = clustering.pad_rows(x, self.store.device)
        else:
            xq = clustering.pad_rows(x, self.store.device)
        ids, dist, parent_info = self._search_device(xq, search_params, tinfo)
        if out_dev.type == "cpu":
            # both results through pinned staging buffers, one synchronisation
            k = int(ids.shape[1])
            stage = self.__dict__.setdefault("_host_stage", {})
            key = (int(ids.shape[0]), k)
            if key not in stage:
                if len(stage) >= _MAX_PLANS:
                    stage.clear()
                stage[key] = (torch.empty(key, dtype=torch.int64).pin_memory(), torch.empty(key, dtype=torch.float32).pin_memory())
            h_ids, h_dist = stage[key]
            n8 = ids.numel() * 8
            if (ids.is_contiguous() and dist.is_contiguous() and ids.dtype == torch.int64
                    and dist.data_ptr() == ids.data_ptr() + n8
                    and ids.untyped_storage().data_ptr() == dist.untyped_storage().data_ptr()):
                # one allocation on the device (_search_ivf), one on the host: a single D2H copy
                if key not in stage.setdefault("_blocks", {}):
                    hb = torch.empty(n8 + dist.numel() * 4, dtype=torch.uint8).pin_memory()
                    stage["_blocks"][key] = {"block": hb, "ids": hb[:n8].view(torch.int64).view(ids.shape),
                                             "dist": hb[n8:].view(torch.float32).view(dist.shape)}
                hb = stage["_blocks"][key]
                dblock = torch.as_strided(ids.view(torch.uint8).reshape(-1), (n8 + dist.numel() * 4,), (1,))
                hb["block"].copy_(dblock, non_blocking=True)
                torch.cuda.current_stream().synchronize()
                res.ids, res.distances = hb["ids"].clone(), hb["dist"].clone()
            else:
                h_ids.copy_(ids, non_blocking=True)
                h_dist.copy_(dist, non_blocking=True)
                torch.cuda.current_stream().synchronize()
                res.ids, res.distances = h_ids.clone(), h_dist.clone()
        else:
            res.ids = ids.to(out_dev, copy=True)
            res.distances = dist.to(out_dev, copy=True)
        tinfo.n_queries = int(x.shape[0])
        tinfo.parent_info = parent_info
        tinfo.total_time_ns = int((time.perf_counter() - t0) * 1e9)
        return res

    def _flat_probe(self, Q: int) -> torch.Tensor:
        """[Q, nlist] probe table of a flat index: every query scans every partition (query_coordinator.cpp:624-626).
        One tensor per batch size for the current store version; captured plans keep their own reference."""
        if self.store.slot_pid.size == 1:
            return None  # flat mode: the scan needs no probe table for a single-list store
        self.store.tables()
        ver = (self.store.uid, self.store.version)
        cache = self.__dict__.setdefault("_flat_probe_cache", {})
        if cache.get("version") != ver:
            cache.clear()
            cache["version"] = ver
        if Q not in cache:
            if len(cache) > 16:
                for kk in [kk for kk in cache if kk != "version"]:
                    del cache[kk]
            slots = torch.tensor([self.store.pid_slot[int(p)] for p in self.store.partition_ids()], dtype=torch.int32,
                                 device=self.store.device)
            cache[Q] = slots[None, :].expand(Q, -1).contiguous()
        return cache[Q]

    def _apply_filter_policy(self) -> None:
        """Filter precision of the partition scan (see FILTER_POLICY): 3xTF32 unless forced to 2; "auto" relaxes a
        two-level index to 2 terms on evidence (_filter_feedback)."""
        if self.store is None:
            return
        two_level = self.parent is not None and self.current_level == 0
        self.store.set_filter_terms(2 if (two_level and FILTER_POLICY == "2") else 3)

    def _filter_feedback(self) -> None:
        mon = self.__dict__.get("_monitor")
        if mon is None or FILTER_POLICY != "auto" or mon.locked:
            return
        got = mon.poll()
        if got is None:
            return
        rescanned, would_fail, queries = got
        if self.store.filter_terms == 2:
            if rescanned > max(2, queries // 100):
                # the 2-term filter cannot separate this data's neighbours often enough: 3xTF32 for good (plans are
                # re-captured: the store version moves)
                self.store.set_filter_terms(3)
                mon.locked = True
                self.filter_fallback = got
        else:
            mon.seen += queries
            mon.would_fail += would_fail
            if mon.seen >= 256:
                if mon.would_fail * 200 <= mon.seen:
                    self.store.set_filter_terms(2)
                mon.seen = mon.would_fail = 0

    def _reset_caches(self) -> None:
        """build() / load() install a new store: captured plans, probe tables and staging buffers of the old one go."""
        self.__dict__.pop("_plans", None)
        self.__dict__.pop("_flat_probe_cache", None)
        self.__dict__.pop("_host_stage", None)
        self.__dict__.pop("_monitor", None)

    def _search_core(self, xq: torch.Tensor, sp: SearchParams):
        """Fixed-nprobe search, launches only (capturable in a CUDA graph): coarse scan -> slot map -> partition scan.
        Returns (ids, distances, probed partition ids or None)."""
        Q = int(xq.shape[0])
        k = int(sp.k) if sp is not None and int(sp.k) > 0 else 1
        if self.parent is None:
            ids, dist = scan_partitions(self.store, xq, self._flat_probe(Q), k, self.metric)
            return ids, dist, None
        if self.parent.parent is None and self.parent.store.slot_pid.size == 1 and Q <= _GRAPH_MAX_Q:
            return self._search_ivf(xq, k, int(sp.nprobe))
        psp = SearchParams()
        psp.batched_scan = True
        psp.k = min(int(sp.nprobe), self.nlist())
        p_ids, _p_dist, _ = self.parent._search_device(xq, psp)
        _, table = self.store.tables()
        slots = torch.empty(p_ids.shape, dtype=torch.int32, device=xq.device)
        check(_lib.load().qk_map_ids_to_slots(ptr(p_ids), p_ids.numel(), ptr(table), table.numel(), ptr(slots), _stream()))
        ids, dist = scan_partitions(self.store, xq, slots, k, self.metric)
        return ids, dist, p_ids

    def _search_ivf(self, xq: torch.Tensor, k: int, nprobe: int, shard_rank: int = 0, shard_world: int = 1):
        """qk_search_ivf: coarse scan of the flat parent -> slot map -> partition scan, one C call, one workspace."""
        lib = _lib.load()
        Q = int(xq.shape[0])
        dev = xq.device
        pstore = self.parent.store
        np_ = max(1, min(int(nprobe), pstore.ntotal))
        pst, _ = pstore.tables(pstore.segment_len(Q, 1))
        st, table = self.store.tables(self.store.segment_len(Q, np_))
        wsb = lib.qk_search_ivf_workspace_bytes(C.byref(pst), C.byref(st), Q, np_, k)
        if wsb == 0:
            check(1)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        # ids and distances share one allocation: search() brings both to the host with ONE copy
        block = torch.empty(Q * k * 12, dtype=torch.uint8, device=dev)
        out_ids = block[: Q * k * 8].view(torch.int64).view(Q, k)
        out_dist = block[Q * k * 8:].view(torch.float32).view(Q, k)
        p_ids = torch.empty((Q, np_), dtype=torch.int64, device=dev)
        # statistics of the partition scan {queries re-scanned exactly, max / total candidates}: the filter precision
        # policy reads them back asynchronously (_FilterMonitor)
        stats = torch.empty(8, dtype=torch.int32, device=dev)
        self._last_stats = stats
        if os.environ.get("QK_SCAN_STATS") == "1":
            global LAST_SCAN_STATS
            LAST_SCAN_STATS = stats
        check(lib.qk_search_ivf(C.byref(pst), C.byref(st), ptr(table), table.numel(), ptr(xq), Q, xq.stride(0), np_,
                                self.metric, k, shard_rank, shard_world, ptr(out_ids), ptr(out_dist), ptr(p_ids), ptr(ws),
                                wsb, ptr(stats), _stream()))
        return out_ids, out_dist, p_ids

    def _plan(self, Q: int, sp: SearchParams) -> "_SearchPlan":
        plans = self.__dict__.setdefault("_plans", {})
        self.store.tables()  # settles store.version
        pv = (-1, -1)
        if self.parent is not None:
            self.parent.store.tables()
            pv = (self.parent.store.uid, self.parent.store.version)
        # store.uid is process-unique: a rebuilt / reloaded index never matches a plan of its previous store
        key = (Q, int(sp.k), int(sp.nprobe), self.metric, self.store.uid, self.store.version, pv)
        plan = plans.get(key)
        if plan is None:
            for old_key in [kk for kk in plans if kk[4:] != key[4:]]:  # the index changed: drop stale plans
                del plans[old_key]
            while len(plans) >= _MAX_PLANS:
                del plans[next(iter(plans))]  # least recently used first (hits are re-inserted below)
            try:
                plan = _SearchPlan(self, Q, sp)
            except RuntimeError:
                # capture failed (e.g. out of memory for the plan's private pool): this batch runs eagerly
                torch.cuda.synchronize()
                return None
            plans[key] = plan
        else:
            plans[key] = plans.pop(key)  # LRU order
        return plan

    def _search_device(self, xq: torch.Tensor, sp: SearchParams, tinfo: SearchTimingInfo | None = None,
                       want_rows: bool = False):
        """Device-resident search: xq [Q, pitch] on the index's device -> (ids, distances, parent timing) device
        tensors; with want_rows (flat index only) the arena rows of the results replace the timing.
        Graph-replayed searches return the plan's static output buffers (valid until the next search)."""
        Q = int(xq.shape[0])
        k = int(sp.k) if sp is not None and int(sp.k) > 0 else 1
        parent_info = SearchTimingInfo()
        if want_rows:
            if self.parent is not None:
                raise RuntimeError("quake_b200: result rows are only available from a flat index")
            return scan_partitions(self.store, xq, self._flat_probe(Q), k, self.metric, want_rows=True)
        use_aps = self.parent is not None and float(sp.recall_target) > 0.0 and not bool(sp.batched_scan)
        if use_aps:
            nlist = self.nlist()
            psp = SearchParams()
            psp.batched_scan = True
            psp.recall_target = sp.recall_target
            psp.k = max(int(nlist * float(sp.initial_search_fraction)), 1)  # query_coordinator.cpp:636-639
            t1 = time.perf_counter()
            p_ids, _p_dist, p_rows = self.parent._search_device(xq, psp, want_rows=True)
            _, table = self.store.tables()
            slots = torch.empty(p_ids.shape, dtype=torch.int32, device=xq.device)
            check(_lib.load().qk_map_ids_to_slots(ptr(p_ids), p_ids.numel(), ptr(table), table.numel(), ptr(slots), _stream()))
            parent_info.total_time_ns = int((time.perf_counter() - t1) * 1e9)
            if self.parent.parent is not None:
                raise RuntimeError("quake_b200: APS needs a flat parent (two-level index)")
            ids, dist, scanned = aps.adaptive_scan(self, xq, p_rows, slots, sp)
            self.last_partitions_scanned = scanned  # per query, device int32 (the reference only reports it for workers)
            if tinfo is not None:
                tinfo.partitions_scanned = int(scanned.sum().item())
            aps_hits = (p_ids, scanned)
        else:
            t1 = time.perf_counter()
            plan = None
            monitored = self.current_level == 0 and self.parent is not None and not _capturing
            if monitored:
                self._filter_feedback()
            if GRAPHS_ENABLED and not _capturing and self.current_level == 0 and Q <= _GRAPH_MAX_Q:
                plan = self._plan(Q, sp)
            if plan is not None:
                ids, dist, p_ids = plan.run(xq)
                if p_ids is not None:
                    p_ids = p_ids.clone()  # the hit window below outlives the plan's static buffer
                stats = plan.stats
            else:
                self._last_stats = None
                ids, dist, p_ids = self._search_core(xq, sp)
                stats = self._last_stats
            if monitored and FILTER_POLICY == "auto":
                self.__dict__.setdefault("_monitor", _FilterMonitor()).submit(stats, Q)
            parent_info.total_time_ns = int((time.perf_counter() - t1) * 1e9)
        if self.parent is not None:
            parent_info.n_queries = Q
            parent_info.n_clusters = self.parent.nlist()
            if self.maintenance_policy is not None and self.current_level == 0 and not _capturing:
                if use_aps:
                    self._record_hits(*aps_hits)
                else:
                    self._record_hits(p_ids)
        return ids, dist, parent_info

    def _record_hits(self, p_ids: torch.Tensor, scanned: torch.Tensor | None = None) -> None:
        """Feed the maintenance policy's hit window (the reference's HitCountTracker, hit_count_tracker.cpp:43-66;
        docs/architecture/architecture.rst:21) with the partitions this batch probed. Stays on the device; only
        maintenance() materialises it."""
        if self.maintenance_policy is not None and p_ids is not None:
            self.maintenance_policy.record_query_hits(p_ids, scanned)

    # ------------------------------------------------------------------ accessors
    def get_ids(self) -> torch.Tensor:
        if self.store is None:
            raise RuntimeError("[QuakeIndex::get_ids()] No partition manager. Index not built?")
        return self.store.all_ids().cpu()  # one gather on the device, one copy

    def get(self, ids: torch.Tensor) -> torch.Tensor:
        if self.store is None:
            raise RuntimeError("[QuakeIndex::get()] No partition manager. Index not built?")
        rows = self.store.find_rows(ids.to(torch.int64))
        if bool((rows < 0).any()):
            raise RuntimeError("ID not found in any partition")
        return self.store.vectors[rows, : self.store.d].to(ids.device)

    def ntotal(self) -> int:
        return self.store.ntotal if self.store is not None else 0

    def nlist(self) -> int:
        return self.store.nlist if self.store is not None else 0

    def d(self) -> int:
        return self.store.d if self.store is not None else 0

    # ------------------------------------------------------------------ add / remove (partition_manager.cpp:123-320)
    def add(self, x: torch.Tensor, ids: torch.Tensor, assignments: torch.Tensor | None = None) -> ModifyTimingInfo:
        if self.store is None:
            raise RuntimeError("[QuakeIndex::add()] No partition manager. Build the index first.")
        info = ModifyTimingInfo()
        s1 = time.perf_counter()
        if x is None or ids is None:
            raise RuntimeError("[PartitionManager] add: vectors or vector_ids is undefined.")
        if x.shape[0] != ids.shape[0]:
            raise RuntimeError("[PartitionManager] add: mismatch in vectors.size(0) and vector_ids.size(0).")
        n = int(x.shape[0])
        info.modify_count = n
        if n == 0:
            return info
        if x.dim() != 2:
            raise RuntimeError("[PartitionManager] add: 'vectors' must be 2D [N, dim].")
        if int(x.shape[1]) != self.store.d:
            raise RuntimeError("[PartitionManager] add: dimension mismatch.")
        dev = self.store.device
        idd = ids.to(device=dev, dtype=torch.int64).contiguous()
        if bool((idd > 2**31 - 1).any()):
            raise RuntimeError("[PartitionManager] add: vector_ids must be less than INT_MAX.")
        if int(torch.unique(idd).numel()) != n:
            raise RuntimeError("[PartitionManager] add: vector_ids must be unique.")
        if assignments is not None and bool((assignments >= self.store.curr_list_id).any()):
            raise RuntimeError("[PartitionManager] add: assignments must be less than partition_store_->curr_list_id_.")
        info.input_validation_time_us = int((time.perf_counter() - s1) * 1e6)
        s2 = time.perf_counter()
        xd = clustering.pad_rows(x, dev)
        _, table = self.store.tables()
        if self.parent is None:
            slots = torch.full((n,), self.store.pid_slot[0] if 0 in self.store.pid_slot else -1, dtype=torch.int32,
                               device=dev)
        else:
            if assignments is not None and assignments.numel() > 0:
                if assignments.shape[0] != n:
                    raise RuntimeError("[PartitionManager] add: assignments.size(0) != vectors.size(0).")
                pid = assignments.to(device=dev, dtype=torch.int64).contiguous()
            else:
                psp = SearchParams()
                psp.k = 1
                psp.nprobe = self.parent.nlist()
                psp.batched_scan = n > 10
                pid, _, _ = self.parent._search_device(xd, psp)
                pid = pid.reshape(-1).contiguous()
            slots = torch.empty((n,), dtype=torch.int32, device=dev)
            check(_lib.load().qk_map_ids_to_slots(ptr(pid), n, ptr(table), table.numel(), ptr(slots), _stream()))
        torch.cuda.synchronize()
        info.find_partition_time_us = int((time.perf_counter() - s2) * 1e6)
        s3 = time.perf_counter()
        if bool((slots < 0).any()):
            raise RuntimeError("List does not exist in add_entries")
        self.store.append(slots, xd, idd)
        torch.cuda.synchronize()
        info.modify_time_us = int((time.perf_counter() - s3) * 1e6)
        return info

    def remove(self, ids: torch.Tensor) -> ModifyTimingInfo:
        if self.store is None:
            raise RuntimeError("[QuakeIndex::remove()] No partition manager. Build the index first.")
        info = ModifyTimingInfo()
        if ids is None or ids.shape[0] == 0:
            return info
        info.modify_count = int(ids.shape[0])
        s = time.perf_counter()
        self.store.remove_ids(ids.to(device=self.store.device, dtype=torch.int64))
        torch.cuda.synchronize()
        info.modify_time_us = int((time.perf_counter() - s) * 1e6)
        return info

    def modify(self, ids: torch.Tensor, x: torch.Tensor) -> ModifyTimingInfo:
        """quake_index.cpp:142-145: remove + add."""
        self.remove(ids)
        return self.add(x, ids)

    # ------------------------------------------------------------------ maintenance
    def initialize_maintenance_policy(self, maintenance_policy_params: MaintenancePolicyParams) -> None:
        """QuakeIndex::initialize_maintenance_policy (quake_index.cpp:148-155)."""
        from .maintenance import MaintenancePolicy
        self.maintenance_policy_params = maintenance_policy_params
        self.maintenance_policy = None
        if self.store is not None and self.parent is not None and self.current_level == 0:
            self.maintenance_policy = MaintenancePolicy(self, maintenance_policy_params)

    # ------------------------------------------------------------------ partition surgery (partition_manager.cpp:393-554)
    def select_partitions(self, partition_ids: torch.Tensor):
        """(centroids [P, d], [vectors_p], [ids_p]) copies of the given partitions (PartitionManager::select_partitions)."""
        pids = [int(p) for p in partition_ids.tolist()]
        cents = self.parent.get(torch.tensor(pids, dtype=torch.int64).to(self.store.device)) if pids else None
        vecs, ids = [], []
        for p in pids:
            v, i = self.store.get_list(p, padded=True)
            vecs.append(v.clone())
            ids.append(i.clone())
        return cents, vecs, ids

    def split_partitions(self, partition_ids: torch.Tensor):
        """PartitionManager::split_partitions (:393-445): every partition is cut in two by k-means (K = 2, the build
        defaults) over its own vectors. Returns (centroids [2P, d], [vectors], [ids]) of the halves, in order."""
        d = self.store.d
        cents_out, vecs_out, ids_out = [], [], []
        for p in [int(x) for x in partition_ids.tolist()]:
            v, i = self.store.get_list(p, padded=True)
            if int(v.shape[0]) < 4:
                raise RuntimeError("Partition must have at least 4 vectors to split.")
            x = v.clone()  # kmeans normalises its input in place for the inner-product metric
            c, counts, offsets, order = clustering.kmeans(x, d, 2, self.metric, 5)
            offs = offsets.cpu().tolist()
            for j in range(2):
                sel = order[offs[j]:offs[j + 1]]
                cents_out.append(c[j, :d])
                vecs_out.append(x.index_select(0, sel))
                ids_out.append(i.index_select(0, sel))
        return torch.stack(cents_out), vecs_out, ids_out

    def add_partitions(self, clustering_triple) -> torch.Tensor:
        """PartitionManager::add_partitions (:490-520): new partition ids curr_list_id .., lists filled, centroids
        added to the parent. Returns the new ids."""
        cents, vecs, ids = clustering_triple
        n = len(vecs)
        st = self.store
        first = int(st.curr_list_id)
        new_pids = torch.arange(first, first + n, dtype=torch.int64)
        for j in range(n):
            cnt = int(vecs[j].shape[0])
            st.add_list(first + j, capacity=(cnt + max(16, cnt // 8) + 3) // 4 * 4)
            st.set_list(first + j, vecs[j][:, : st.d], ids[j])
        self.parent.add(cents[:, : st.d].contiguous(), new_pids)
        return new_pids

    def delete_partitions(self, partition_ids: torch.Tensor, reassign: bool = False) -> None:
        """PartitionManager::delete_partitions (:522-554): the centroids leave the parent, the lists leave the store;
        with `reassign` their vectors are added back (each to its nearest remaining partition)."""
        if self.parent is None:
            raise RuntimeError("Index is not partitioned")
        _, vecs, ids = self.select_partitions(partition_ids)
        self.parent.remove(partition_ids)
        for p in partition_ids.tolist():
            self.store.remove_list(int(p))
        self.store.maybe_compact()
        if reassign:
            keep = [(v, i) for v, i in zip(vecs, ids) if int(v.shape[0]) > 0]
            if keep:
                self.add(torch.cat([v[:, : self.store.d] for v, _ in keep]), torch.cat([i for _, i in keep]))

    def refine_partitions(self, partition_ids: torch.Tensor | None = None, iterations: int = 0) -> None:
        """PartitionManager::refine_partitions (partition_manager.cpp:447-488): Lloyd refit restricted to
        the given partitions; centroids in the parent are replaced (remove + add)."""
        if self.parent is None:
            raise RuntimeError("Index is not partitioned")
        if partition_ids is None:
            partition_ids = self.parent.get_ids()
        pids = [int(p) for p in partition_ids.tolist()]
        if not pids:
            return
        dev = self.store.device
        d = self.store.d
        pid_t = torch.tensor(pids, dtype=torch.int64)
        cents = clustering.pad_rows(self.parent.get(pid_t.to(dev)), dev)
        rows = self.store.rows_of(pids)
        allv = self.store.vectors.index_select(0, rows)
        alli = self.store.ids.index_select(0, rows)
        new_c, counts, nv, ni = clustering.kmeans_refine(cents, d, allv, alli, self.metric, int(iterations))
        self.store.replace_lists(pids, counts.cpu().numpy(), nv, ni)
        self.parent.modify(pid_t, new_c[:, :d])

    def maintenance(self) -> MaintenanceTimingInfo:
        """QuakeIndex::maintenance (quake_index.cpp:157-163) -> MaintenancePolicy::perform_maintenance
        (maintenance_policies.cpp:33-172): nothing happens until the hit window is full; then hit rates feed the cost
        model (with the scan latency measured on this GPU), partitions are deleted / split accordingly and the
        neighbourhood of the new partitions is refit (quake_b200/maintenance.py)."""
        if self.maintenance_policy_params is None:
            raise RuntimeError("[QuakeIndex::maintenance()] No maintenance policy set.")
        if self.maintenance_policy is None:
            return MaintenanceTimingInfo()  # flat index: nothing to maintain
        return self.maintenance_policy.perform_maintenance()

    # ------------------------------------------------------------------ save / load (quake_index.cpp:170-267)
    def save(self, dir_path: str) -> None:
        if os.path.exists(dir_path) and not os.path.isdir(dir_path):
            raise RuntimeError("save path exists but is not a directory: " + dir_path)
        os.makedirs(dir_path, exist_ok=True)
        with open(os.path.join(dir_path, "metadata.txt"), "w") as f:
            f.write(f"metric={self.metric}\nlevel={self.current_level}\nntotal={self.ntotal()}\nnlist={self.nlist()}\n")
        self._save_partitions(os.path.join(dir_path, "partitions"))
        if self.parent is not None:
            self.parent.save(os.path.join(dir_path, "parent"))

    def _save_partitions(self, path: str) -> None:
        """DynamicInvertedLists::save, format v3 (dynamic_inverted_list.cpp:338-419): 32-byte header
        {magic u32, version u32, nlist u64, code_size u64, num_partitions u64}, offsets u64[np+1],
        partition ids u64[np], then per partition codes[n x code_size] followed by ids[n x 8]."""
        st = self.store
        pids = st.partition_ids()
        code_size = st.d * 4
        sizes = np.array([st.size_of(int(p)) for p in pids], dtype=np.uint64)
        chunk = sizes * np.uint64(code_size + 8)
        offsets = np.concatenate([[0], np.cumsum(chunk)]).astype(np.uint64)
        # one gather of the live rows on the device, one copy each for vectors and ids
        rows = st.rows_of(pids) if len(pids) else torch.zeros(0, dtype=torch.int64, device=st.device)
        vec_h = st.vectors.index_select(0, rows)[:, : st.d].contiguous().cpu().numpy()
        ids_h = st.ids.index_select(0, rows).cpu().numpy()
        starts = np.concatenate([[0], np.cumsum(sizes.astype(np.int64))])
        with open(path, "wb") as f:
            f.write(struct.pack("<IIQQQ", SERIALIZATION_MAGIC, SERIALIZATION_VERSION, len(pids), code_size, len(pids)))
            f.write(offsets.tobytes())
            f.write(pids.astype(np.uint64).tobytes())
            for j in range(len(pids)):
                a, b = int(starts[j]), int(starts[j + 1])
                f.write(vec_h[a:b].tobytes())
                f.write(ids_h[a:b].tobytes())

    def load(self, dir_path: str, n_workers: int = 0) -> None:
        if not os.path.isdir(dir_path):
            raise RuntimeError("Cannot load QuakeIndex, directory does not exist: " + dir_path)
        dev = _device()
        with open(os.path.join(dir_path, "metadata.txt")) as f:
            for line in f:
                if "=" not in line:
                    continue
                key, val = line.strip().split("=", 1)
                if key == "metric":
                    self.metric = int(val)
                elif key == "level":
                    self.current_level = int(val)
        self._load_partitions(os.path.join(dir_path, "partitions"), dev)
        self._reset_caches()
        pdir = os.path.join(dir_path, "parent")
        if os.path.isdir(pdir):
            self.parent = QuakeIndex()
            self.parent.load(pdir, n_workers)
        else:
            self.parent = None
        self.initialize_maintenance_policy(MaintenancePolicyParams())
        self._apply_filter_policy()

    def _load_partitions(self, path: str, dev) -> None:
        with open(path, "rb") as f:
            raw = f.read()
        magic, version, _nlist, code_size, nparts = struct.unpack_from("<IIQQQ", raw, 0)
        if magic != SERIALIZATION_MAGIC:
            raise RuntimeError("Invalid file format (bad magic number).")
        if version != SERIALIZATION_VERSION:
            raise RuntimeError("Unsupported file version: " + str(version))
        d = code_size // 4
        offsets = np.frombuffer(raw, dtype=np.uint64, count=nparts + 1, offset=32)
        pids = np.frombuffer(raw, dtype=np.uint64, count=nparts, offset=32 + 8 * (nparts + 1)).astype(np.int64)
        start = 32 + 8 * (nparts + 1) + 8 * nparts
        rec = code_size + 8
        vec_parts, id_parts, counts = [], [], []
        for i in range(nparts):
            size = int(offsets[i + 1] - offsets[i])
            if size % rec:
                raise RuntimeError("Partition chunk size not divisible by (code_size+sizeof(idx_t))")
            nv = size // rec
            o = start + int(offsets[i])
            vec_parts.append(np.frombuffer(raw, dtype=np.float32, count=nv * d, offset=o).reshape(nv, d))
            id_parts.append(np.frombuffer(raw, dtype=np.int64, count=nv, offset=o + nv * code_size))
            counts.append(nv)
        self.store = PartitionStore(d, dev)
        vecs = np.concatenate(vec_parts) if vec_parts else np.zeros((0, d), np.float32)
        idv = np.concatenate(id_parts) if id_parts else np.zeros((0,), np.int64)
        xd = clustering.pad_rows(torch.from_numpy(vecs.copy()), dev)
        idd = torch.from_numpy(idv.copy()).to(dev)
        self.store.init_from_sorted(xd, idd, None, np.array(counts, dtype=np.int64), pids)

    def __repr__(self):
        return '{"current_level": %d, }' % self.current_level
