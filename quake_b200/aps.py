"""Adaptive Partition Scanning (APS) driver: recall_target > 0 on the serial-scan path.

Host-side mirror of the APS part of QueryCoordinator::serial_scan
(/root/reference/src/cpp/src/query_coordinator.cpp:521-579). The arithmetic runs on the device
(qk_aps_boundary_distances, qk_scan_partitions, qk_aps_advance -- csrc/aps.cu); this file sequences the rounds.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr

_FIRST_ROUND = 2   # probe ranks scanned in the first round (pseudo-queries with a full top-k each)
_MAX_ROUND = 128   # ... doubling every round while most queries are still active
_MAX_PSEUDO = 1 << 17  # pseudo-queries (query, rank) per scan call
_MAX_COLLECT_ENTRIES = 1 << 26  # (query, rank, k) result entries of one collect round (12 B each: 768 MB)

_beta_tables: dict = {}
_TRACE = os.environ.get("QK_APS_TRACE") == "1"
_trace_t = [0.0]


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def beta_table(d: int, device) -> torch.Tensor:
    """incomplete_beta_lookup's table (geometry.h:163-186) for dimension d, computed on the host in double and
    cached on the device. (The reference keeps ONE process-global table initialised with the first d it sees;
    here every dimension gets its own.)"""
    key = (int(d), str(device))
    if key not in _beta_tables:
        arr = np.empty(1001, dtype=np.float64)
        check(_lib.load().qk_host_beta_table(int(d), arr.ctypes.data_as(C.POINTER(C.c_double))))
        _beta_tables[key] = torch.from_numpy(arr).to(device)
    return _beta_tables[key]


def adaptive_scan(index, xq: torch.Tensor, cand_rows: torch.Tensor, slots: torch.Tensor, sp):
    """xq [Q, pitch] device queries; cand_rows [Q, m] arena rows of the rank-ordered candidate centroids in the
    parent store; slots [Q, m] list slots of the candidates in this index's store (-1 = skip).
    Returns (ids [Q, k], distances [Q, k], partitions scanned per query [Q])."""
    from .index import scan_partitions  # local import: index.py imports this module

    lib = _lib.load()
    Q, m = int(slots.shape[0]), int(slots.shape[1])
    if m < 2:
        # compute_recall_profile (geometry.h:350-352)
        raise RuntimeError("Boundary distances must have at least 2 partitions to create an estimate.")
    k = max(int(sp.k), 1)
    if k > 1024:
        raise ValueError("quake_b200: APS supports k <= 1024")
    dev = xq.device
    store, pstore = index.store, index.parent.store
    d, metric = store.d, index.metric
    ip = metric == _lib.QK_METRIC_INNER_PRODUCT
    slots = slots.to(torch.int32).contiguous()
    cand_rows = cand_rows.to(torch.int64).contiguous()

    boundary = torch.empty((Q, m), dtype=torch.float32, device=dev)
    check(lib.qk_aps_boundary_distances(ptr(xq), Q, xq.stride(0), d, ptr(pstore.vectors), pstore.pitch, ptr(cand_rows),
                                        m, metric, ptr(boundary), _stream()))
    table = beta_table(d, dev) if (bool(sp.use_precomputed) and not ip) else None

    run_ids = torch.full((Q, k), -1, dtype=torch.int64, device=dev)
    run_dist = torch.full((Q, k), float("-inf") if ip else float("inf"), dtype=torch.float32, device=dev)
    run_cnt = torch.zeros(Q, dtype=torch.int32, device=dev)
    radius = torch.full((Q,), -1000000.0 if ip else 1000000.0, dtype=torch.float32, device=dev)  # query_coordinator.cpp:524-527
    have = torch.zeros(Q, dtype=torch.int32, device=dev)
    probs = torch.zeros((Q, m), dtype=torch.float32, device=dev)
    done = torch.zeros(Q, dtype=torch.int32, device=dev)
    scanned = torch.zeros(Q, dtype=torch.int32, device=dev)
    still = torch.zeros(1, dtype=torch.int32, device=dev)

    # Rounds of R probe ranks. While some active query holds fewer than k results (no finite k-th distance yet) a round
    # scans every (query, rank) as its own pseudo-query with a full top-k (the first round, normally). After that a
    # round is ONE scan per active query in collect mode: the query's current k-th distance -- a valid bound for
    # everything that can still enter its top-k -- is turned into a fixed filter threshold, all rows under it are
    # refined exactly and grouped by rank, and the advance kernel replays the reference's list-by-list loop on them.
    # A list of 610 rows then costs a handful of exact distances instead of a 100-deep top-k of its own.
    active = torch.arange(Q, dtype=torch.int32, device=dev)
    if _TRACE:
        import time as _t
        torch.cuda.synchronize()
        _trace_t[0] = _t.perf_counter()
    p, R = 0, _FIRST_ROUND
    collect_ok = os.environ.get("QK_APS_COLLECT", "1") != "0" and store.list_size.size > 0 and \
        int(store.list_size.max()) <= _lib.QK_SEGMENT_ROWS
    first = True
    while p < m and active.numel() > 0:
        Qa = int(active.numel())
        act64 = active.to(torch.int64)
        xa = xq.index_select(0, act64)
        use_collect = collect_ok and not first and int(run_cnt.index_select(0, act64).min().item()) >= k
        done_round = False
        if use_collect:
            r_eff = min(R, m - p)
            probe = slots.index_select(0, act64)[:, p:p + r_eff].contiguous()
            st, _ = store.tables(_lib.QK_SEGMENT_ROWS)  # whole lists: one segment each (pair slot == probe rank)
            st3 = _lib.QkStore.from_buffer_copy(st)
            st3.filter_terms = 3  # no re-scan monitor on this path: the tight filter
            thr = torch.empty(Qa, dtype=torch.int32, device=dev)
            check(lib.qk_aps_thresholds(ptr(active), Qa, ptr(xq), xq.stride(0), d, ptr(run_dist), ptr(run_cnt), k, metric,
                                        float(store.max_row_norm), 3, ptr(thr), _stream()))
            wsb = lib.qk_scan_workspace_bytes(C.byref(st3), Qa, r_eff, k)
            ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
            r_ids = torch.empty((Qa, r_eff, k), dtype=torch.int64, device=dev)
            r_dist = torch.empty((Qa, r_eff, k), dtype=torch.float32, device=dev)
            r_cnt = torch.empty((Qa, r_eff), dtype=torch.int32, device=dev)
            ovf = torch.empty(Qa, dtype=torch.int32, device=dev)
            check(lib.qk_scan_collect(C.byref(st3), ptr(xa), Qa, xa.stride(0), ptr(probe), r_eff, metric, k, ptr(thr),
                                      ptr(r_ids), ptr(r_dist), ptr(r_cnt), ptr(ovf), ptr(ws), wsb, _stream()))
            if int(ovf.max().item()) == 0:  # (one host read; an overflow sends the round down the pseudo-query path)
                check(lib.qk_aps_advance(ptr(active), Qa, r_eff, p, m, k, d, metric, ptr(slots), ptr(r_ids), ptr(r_dist),
                                         ptr(r_cnt), ptr(boundary), ptr(table), float(sp.recall_target),
                                         float(sp.recompute_threshold), int(bool(sp.use_precomputed)), ptr(run_ids),
                                         ptr(run_dist), ptr(run_cnt), ptr(radius), ptr(have), ptr(probs), ptr(done),
                                         ptr(scanned), ptr(still), _stream()))
                done_round = True
        if not done_round:
            r_eff = min(R if not first else _FIRST_ROUND, m - p, max(1, _MAX_PSEUDO // Qa))
            xrep = xa.repeat_interleave(r_eff, dim=0)                       # pseudo-query (a, r) = a * r_eff + r
            probe = slots.index_select(0, act64)[:, p:p + r_eff].reshape(-1, 1).contiguous()
            r_ids, r_dist = scan_partitions(store, xrep, probe, k, metric, filter_terms=3)
            check(lib.qk_aps_advance(ptr(active), Qa, r_eff, p, m, k, d, metric, ptr(slots), ptr(r_ids), ptr(r_dist), None,
                                     ptr(boundary), ptr(table), float(sp.recall_target), float(sp.recompute_threshold),
                                     int(bool(sp.use_precomputed)), ptr(run_ids), ptr(run_dist), ptr(run_cnt), ptr(radius),
                                     ptr(have), ptr(probs), ptr(done), ptr(scanned), ptr(still), _stream()))
        first = False
        p += r_eff
        n_still = int(still.item())   # one host read per round: how many queries go on
        if _TRACE:
            import time as _t
            torch.cuda.synchronize()
            now = _t.perf_counter()
            print(f"[aps] ranks {p - r_eff}..{p - 1} {'collect' if done_round else 'pseudo'} active {Qa} -> {n_still} "
                  f"({(now - _trace_t[0]) * 1e3:.2f} ms)", flush=True)
            _trace_t[0] = now
        if n_still == 0:
            break
        # every round streams the probed lists again: while most queries are still going, bigger rounds cost less than
        # the ranks a finishing query over-scans; once most have stopped, keep the round size. Once a round touches most
        # of the index anyway (active queries x ranks >= half the lists: a pass over all of HBM whatever the round
        # size), the passes are what costs, not the ranks: the scanned prefix then grows five-fold per round (C3:
        # rounds 14..69 and 70..326 instead of five doubling rounds, each a full pass over the 5 GB index).
        if 2 * n_still > Qa:
            R = min(2 * R, _MAX_ROUND)
            if collect_ok and 2 * n_still * R >= store.nlist:
                dense_r = min(4 * p, max(1, _MAX_COLLECT_ENTRIES // (n_still * k)))
                R = max(R, dense_r)
        active = torch.nonzero(done == 0).reshape(-1).to(torch.int32)
    return run_ids, run_dist, scanned
