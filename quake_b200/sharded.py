"""Inverted lists sharded over the GPUs of one box (BASELINE.json config 4).

One process per GPU (torch.distributed). Partition p lives on rank ``p % world`` -- the round-robin of the
reference's own partition-to-worker distribution (/root/reference/src/cpp/src/partition_manager.cpp:599-602).
The centroid (parent) index is replicated: every rank runs the coarse scan for all queries (redundant, cheap),
scans the probed lists it owns, and the per-rank partial top-k lists (Q * k * 12 bytes per rank) are exchanged and
merged on every rank, exactly what the reference does per core with TopkBuffer::batch_add
(src/cpp/src/query_coordinator.cpp:167-173). There is no other data-path collective.

The exchange is ONE kernel per rank over NVLink peer memory (qk_exchange_merge_topk, csrc/exchange.cu): every rank
pushes its partial into every peer's buffer with remote stores, signals with one flag per CTA, waits for the peers'
pushes and merges -- no NCCL call on the query path (NCCL is used once, at set-up, to distribute the CUDA IPC handles
of the buffers). The per-rank step (coarse scan -> owned-list scan -> exchange + merge) is captured in a CUDA graph
per (batch size, k, nprobe). QK_EXCHANGE=nccl selects the plain path (two all-gathers + qk_merge_topk) instead.

The result is bit-identical to the unsharded index: a query's answer is the top-k of the union of per-list
top-k's, and the merge orders by (distance, id) like the single-GPU refine step.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, clustering
from ._lib import check, ptr
from .params import IndexBuildParams, SearchParams, SearchResult, SearchTimingInfo, str_to_metric
from .store import PartitionStore


def owner_of(partition_ids, world: int):
    """Rank that owns each partition id."""
    return partition_ids % world


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def merge_partials_device(part_dist: torch.Tensor, part_ids: torch.Tensor, k: int, metric: int):
    """[S, Q, k] partial results on the device -> merged ([Q, k] ids, [Q, k] distances) via qk_merge_topk."""
    lib = _lib.load()
    S, Q = int(part_ids.shape[0]), int(part_ids.shape[1])
    dev = part_ids.device
    out_ids = torch.empty((Q, k), dtype=torch.int64, device=dev)
    out_dist = torch.empty((Q, k), dtype=torch.float32, device=dev)
    check(lib.qk_merge_topk(ptr(part_dist.contiguous()), ptr(part_ids.contiguous()), S, Q, k, metric, ptr(out_ids),
                            ptr(out_dist), _stream()))
    return out_ids, out_dist


def gather_and_merge(ids: torch.Tensor, distances: torch.Tensor, k: int, metric: int, group=None, merge=None):
    """All-gather the [Q, k] partial top-k of every rank and merge them (same result on every rank).
    `merge(part_dist [S,Q,k], part_ids [S,Q,k], k, metric)` defaults to the device kernel; the CPU (gloo) tests
    of this plumbing inject a host merge, the product path never does."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return ids, distances
    ids = ids.contiguous()
    distances = distances.contiguous()
    Q = int(ids.shape[0])
    all_ids = torch.empty((world * Q, ids.shape[1]), dtype=ids.dtype, device=ids.device)
    all_dist = torch.empty((world * Q, distances.shape[1]), dtype=distances.dtype, device=distances.device)
    dist.all_gather_into_tensor(all_ids, ids, group=group)      # rank r's block at rows [r * Q, (r + 1) * Q)
    dist.all_gather_into_tensor(all_dist, distances, group=group)
    return (merge or merge_partials_device)(all_dist.view(world, Q, -1), all_ids.view(world, Q, -1), k, metric)


def mask_foreign_probes(partition_ids: torch.Tensor, slots: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Probe slots of the lists this rank does not own become -1 (skipped by the scan, query_coordinator.cpp:540)."""
    mine = owner_of(partition_ids, world) == rank
    return torch.where(mine & (partition_ids >= 0), slots, torch.full_like(slots, -1))


class _ShardPlan:
    """One captured per-rank step (coarse scan -> owned-list scan -> peer exchange + merge). Every rank captures and
    replays in lockstep: the warm-up runs and every replay are collective exchanges."""

    def __init__(self, sh: "ShardedQuakeIndex", xq: torch.Tensor, sp: SearchParams):
        dev = xq.device
        self.xq = torch.zeros_like(xq)
        self.xq.copy_(xq)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                sh._step(self.xq, sp)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self.ids, self.dist = sh._step(self.xq, sp)
        self.keep = (sh.local.store.tables_snapshot(), sh.local.parent.store.tables_snapshot())

    def run(self, xq: torch.Tensor):
        if xq.data_ptr() != self.xq.data_ptr():
            self.xq.copy_(xq, non_blocking=True)
        self.graph.replay()
        return self.ids, self.dist


class ShardedQuakeIndex:
    """A two-level QuakeIndex whose lists are sharded over the ranks of `group`.

    build()/search() are collective calls: every rank passes the same queries and gets the same result."""

    def __init__(self, group=None, rank: int | None = None, world: int | None = None):
        """rank/world default to the process group's; passing them explicitly builds one shard of a W-way split
        inside a single process (tests; search_partial() then yields that shard's un-merged answer)."""
        from .index import QuakeIndex
        self.group = group
        self.rank = rank if rank is not None else (dist.get_rank(group) if dist.is_initialized() else 0)
        self.world = world if world is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
        self.local = QuakeIndex()  # parent replicated; store holds the owned lists only
        self.metric = 1
        self._ntotal = 0
        self._peer = {}    # (Q, k) -> (local buffer ptr, ctypes array of world peer pointers)
        self._plans = {}   # (Q, k, nprobe, store versions) -> captured per-rank step
        self._peer_failed = None

    # ------------------------------------------------------------------ construction
    def shard_from(self, full) -> None:
        """Keep the lists of a replicated, fully built index that this rank owns (tests / small indexes)."""
        self.metric = full.metric
        self.local.metric = full.metric
        self.local.parent = full.parent
        self.local.build_params = full.build_params
        self.local.maintenance_policy_params = None
        st = full.store
        pids = st.partition_ids()
        mine = pids[owner_of(pids, self.world) == self.rank]
        new = PartitionStore(st.d, st.device)
        vecs, ids, counts = [], [], []
        for p in mine:
            v, i = st.get_list(int(p), padded=True)
            vecs.append(v)
            ids.append(i)
            counts.append(int(v.shape[0]))
        if vecs:
            new.init_from_sorted(torch.cat(vecs).contiguous(), torch.cat(ids).contiguous(), None,
                                 np.asarray(counts, dtype=np.int64), mine)
        new.curr_list_id = st.curr_list_id
        self.local.store = new
        self._ntotal = full.ntotal()

    def build(self, x_local: torch.Tensor, ids_local: torch.Tensor, build_params: IndexBuildParams) -> None:
        """Distributed build from per-rank slices of the data (config 4 is generated per shard).

        1. every rank contributes its share of the k-means training sample (all-gather) and trains the SAME
           centroids (deterministic kernels, identical input => identical centroids on every rank);
        2. every rank assigns its own slice (the k-means assign kernel);
        3. vectors travel to the rank that owns their partition (one all-to-all);
        4. every rank lays out its lists."""
        from .index import QuakeIndex, _device
        dev = _device()
        self.metric = str_to_metric(build_params.metric)
        K = int(build_params.nlist)
        d = int(x_local.shape[1])
        xd = clustering.pad_rows(x_local, dev)
        if xd.data_ptr() == x_local.data_ptr():
            xd = xd.clone()
        idd = ids_local.to(device=dev, dtype=torch.int64).contiguous()
        if self.metric == _lib.QK_METRIC_INNER_PRODUCT:
            check(_lib.load().qk_normalize_rows(ptr(xd), int(xd.shape[0]), xd.stride(0), d, _stream()))
        # ---- 1. training sample
        n_local = int(xd.shape[0])
        want = (K * clustering.MAX_POINTS_PER_CENTROID + self.world - 1) // self.world
        take = min(n_local, want)
        perm = clustering.rand_perm_prefix(n_local, clustering.FAISS_SEED + self.rank, take)
        sample = xd[torch.from_numpy(perm).to(dev)].contiguous()
        if self.world > 1:
            sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(self.world)]
            dist.all_gather(sizes, torch.tensor([take], dtype=torch.int64, device=dev), group=self.group)
            sizes = [int(s.item()) for s in sizes]
            parts = [torch.empty((s, sample.shape[1]), dtype=torch.float32, device=dev) for s in sizes]
            dist.all_gather(parts, sample, group=self.group)
            sample = torch.cat(parts).contiguous()
        trained = clustering.train_centroids(sample, d, K, self.metric, int(build_params.niter))
        centroids = trained
        if self.metric == _lib.QK_METRIC_INNER_PRODUCT:
            centroids = trained.clone()
            check(_lib.load().qk_normalize_rows(ptr(centroids), K, centroids.stride(0), d, _stream()))
        # ---- 2. assignment of the local slice (against the un-normalised trained centroids, like kmeans())
        assign = clustering.assign_points(xd, d, trained, self.metric, filt=clustering.AssignFilter(dev)).to(torch.int64)
        # ---- 3. exchange: rows sorted by (owner, partition), one all-to-all for vectors, ids, partition ids
        owner = owner_of(assign, self.world)
        order = torch.argsort(owner * K + assign, stable=True)
        send_counts = torch.bincount(owner, minlength=self.world)
        xs, ids_s, pid_s = xd[order].contiguous(), idd[order].contiguous(), assign[order].contiguous()
        if self.world > 1:
            recv_counts = torch.empty_like(send_counts)
            dist.all_to_all_single(recv_counts, send_counts, group=self.group)
            sc, rc = send_counts.tolist(), recv_counts.tolist()
            n_recv = int(sum(rc))
            xr = torch.empty((n_recv, xs.shape[1]), dtype=torch.float32, device=dev)
            ir = torch.empty(n_recv, dtype=torch.int64, device=dev)
            pr = torch.empty(n_recv, dtype=torch.int64, device=dev)
            dist.all_to_all_single(xr, xs, rc, sc, group=self.group)
            dist.all_to_all_single(ir, ids_s, rc, sc, group=self.group)
            dist.all_to_all_single(pr, pid_s, rc, sc, group=self.group)
        else:
            xr, ir, pr = xs, ids_s, pid_s
        # ---- 4. local lists (grouped by partition id; arrival order inside a list: by source rank, then input order)
        mine = np.arange(self.rank, K, self.world, dtype=np.int64)
        local_slot = (pr - self.rank) // self.world
        order2 = torch.argsort(local_slot, stable=True)
        counts = torch.bincount(local_slot, minlength=len(mine)).cpu().numpy()
        self.local.metric = self.metric
        self.local.build_params = build_params
        self.local.store = PartitionStore(d, dev)
        self.local.store.init_from_sorted(xr, ir, order2, counts, mine)
        self.local.store.curr_list_id = K
        parent = QuakeIndex(1)
        pp = IndexBuildParams()
        pp.metric = build_params.metric
        parent.build(centroids[:, :d], torch.arange(K, dtype=torch.int64), pp)
        self.local.parent = parent
        self.local.maintenance_policy_params = None
        total = torch.tensor([int(xr.shape[0])], dtype=torch.int64, device=dev)
        if self.world > 1:
            dist.all_reduce(total, group=self.group)
        self._ntotal = int(total.item())

    # ------------------------------------------------------------------ queries
    def ntotal(self) -> int:
        return self._ntotal

    def nlist(self) -> int:
        # the number of partitions of the whole index = the number of centroids the (flat) parent holds
        return self.local.parent.ntotal() if self.local.parent is not None else 0

    # ------------------------------------------------------------------ exchange over NVLink peer memory
    def _use_peer_exchange(self) -> bool:
        return (self.world > 1 and dist.is_initialized() and self._peer_failed is None
                and os.environ.get("QK_EXCHANGE", "peer") != "nccl")

    def exchange_kind(self) -> str:
        if self.world <= 1:
            return "none (one rank)"
        if self._use_peer_exchange():
            return "one kernel per rank over NVLink peer memory (remote stores + flags + merge); no NCCL call"
        why = f" ({self._peer_failed})" if self._peer_failed else ""
        return "NCCL all_gather x2 + qk_merge_topk" + why

    def _peer_buffers(self, Q: int, k: int):
        """Collective, once per (Q, k): allocate this rank's peer buffer, all-gather the CUDA IPC handles, open the
        peers' buffers. Returns the ctypes array qk_exchange_merge_topk takes."""
        key = (int(Q), int(k))
        if key in self._peer:
            return self._peer[key][1]
        lib = _lib.load()
        dev = self.local.store.device
        nbytes = lib.qk_peer_buffer_bytes(Q, k, self.world)
        local = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        rc = lib.qk_peer_alloc(nbytes, C.byref(local), handle)
        ok = torch.tensor([1 if rc == 0 else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        if int(ok.item()) == 0:
            raise _lib.QuakeB200Error("peer buffer allocation failed: " + lib.qk_last_error().decode())
        mine = torch.tensor(list(handle), dtype=torch.uint8, device=dev)
        allh = torch.empty((self.world, 64), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(allh, mine, group=self.group)
        allh = allh.cpu().numpy()
        ptrs = (C.c_void_p * self.world)()
        opened = 1
        for r in range(self.world):
            if r == self.rank:
                ptrs[r] = local.value
                continue
            h = (C.c_ubyte * 64)(*allh[r].tolist())
            p = C.c_void_p()
            if lib.qk_peer_open(h, C.byref(p)) != 0:
                opened = 0
                break
            ptrs[r] = p.value
        ok = torch.tensor([opened], dtype=torch.int32, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        if int(ok.item()) == 0:
            raise _lib.QuakeB200Error("peer buffer could not be opened on every rank: " + lib.qk_last_error().decode())
        self._peer[key] = (local, ptrs)
        return ptrs

    def exchange_and_merge(self, ids: torch.Tensor, distances: torch.Tensor, k: int):
        """Partial [Q, k] results of this rank -> merged result (the same on every rank). Collective."""
        if self.world <= 1:
            return ids, distances
        if self._use_peer_exchange():
            try:
                ptrs = self._peer_buffers(int(ids.shape[0]), k)
            except _lib.QuakeB200Error as e:  # no peer access on this box: every rank takes the NCCL path from now on
                self._peer_failed = str(e)
                ptrs = None
            if ptrs is not None:
                lib = _lib.load()
                ids, distances = ids.contiguous(), distances.contiguous()
                out_ids = torch.empty_like(ids)
                out_dist = torch.empty_like(distances)
                check(lib.qk_exchange_merge_topk(ptr(ids), ptr(distances), int(ids.shape[0]), k, self.metric, self.rank,
                                                 self.world, ptrs, ptr(out_ids), ptr(out_dist), _stream()))
                return out_ids, out_dist
        return gather_and_merge(ids, distances, k, self.metric, self.group)

    def _step(self, xq: torch.Tensor, sp: SearchParams):
        ids, dd = self.search_partial(xq, sp)
        return self.exchange_and_merge(ids, dd, max(int(sp.k), 1))

    def search_device(self, xq: torch.Tensor, sp: SearchParams):
        """xq [Q, pitch] on this rank's device (the same queries on every rank) -> merged (ids, distances).
        The per-rank step is replayed from a CUDA graph when the exchange runs on peer memory (an NCCL collective
        inside a captured graph is avoided on purpose)."""
        from . import index as _qi
        Q = int(xq.shape[0])
        if not (_qi.GRAPHS_ENABLED and self._use_peer_exchange() and Q <= _qi._GRAPH_MAX_Q):
            return self._step(xq, sp)
        st = self.local.store
        st.tables()
        self.local.parent.store.tables()
        key = (Q, int(sp.k), int(sp.nprobe), st.uid, st.version, self.local.parent.store.uid, self.local.parent.store.version)
        plan = self._plans.get(key)
        if plan is None:
            self._peer_buffers(Q, max(int(sp.k), 1))  # collective set-up stays outside the capture
            if not self._use_peer_exchange():
                return self._step(xq, sp)
            for kk in [kk for kk in self._plans if kk[3:] != key[3:]]:
                del self._plans[kk]
            plan = self._plans[key] = _ShardPlan(self, xq, sp)
        return plan.run(xq)

    def search_partial(self, xq: torch.Tensor, sp: SearchParams):
        """This rank's partial top-k: coarse scan for all queries, partition scan of the probed lists it owns."""
        from .index import scan_partitions
        idx = self.local
        k = max(int(sp.k), 1)
        if idx.parent.parent is None and idx.parent.store.slot_pid.size == 1 and idx.store.slot_pid.size > 0:
            # one C call: coarse scan -> id map + ownership filter inside the pair expansion -> scan of the owned lists
            ids, dd, _ = idx._search_ivf(xq, k, int(sp.nprobe), self.rank, self.world)
            return ids, dd
        psp = SearchParams()
        psp.batched_scan = True
        psp.k = min(int(sp.nprobe), self.nlist())
        p_ids, _, _ = idx.parent._search_device(xq, psp)
        _, table = idx.store.tables()
        slots = torch.empty(p_ids.shape, dtype=torch.int32, device=xq.device)
        check(_lib.load().qk_map_ids_to_slots(ptr(p_ids), p_ids.numel(), ptr(table), table.numel(), ptr(slots), _stream()))
        slots = mask_foreign_probes(p_ids, slots, self.rank, self.world)
        return scan_partitions(idx.store, xq, slots, k, self.metric)

    def search(self, x: torch.Tensor, search_params: SearchParams) -> SearchResult:
        res = SearchResult()
        res.timing_info = SearchTimingInfo()
        res.timing_info.search_params = search_params
        if x is None or x.numel() == 0:
            res.ids = torch.empty((0,), dtype=torch.int64)
            res.distances = torch.empty((0,), dtype=torch.float32)
            return res
        if float(search_params.recall_target) > 0.0:
            raise RuntimeError("quake_b200: APS on a sharded index is not supported (fixed nprobe only)")
        xq = clustering.pad_rows(x, self.local.store.device)
        ids, dd = self.search_device(xq, search_params)
        res.ids, res.distances = ids.to(x.device), dd.to(x.device)
        res.timing_info.n_queries = int(x.shape[0])
        return res
