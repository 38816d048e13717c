"""Device-resident dynamic partition store.

B200 counterpart of the reference's ``faiss::DynamicInvertedLists`` + ``IndexPartition``
(/root/reference/src/cpp/include/dynamic_inverted_list.h:33, include/index_partition.h:24-29,
src/index_partition.cpp:52-102, 247-255): every partition ("list") owns a contiguous run of rows
``[row0, row0 + cap)`` of one HBM arena (vectors ``[rows, pitch]`` float32 + ids ``[rows]`` int64), of
which the first ``size`` are live. Appends fill the slack; a list that outgrows its capacity is moved
to the end of the arena with doubled capacity (the reference doubles per-list mallocs); removal is the
reference's swap-with-last. Lists are cut into scan segments of <= QK_SEGMENT_ROWS rows.

The list table (row0 / size / capacity per list) is mirrored on the host for planning, as in the reference; the
payload, the lookup tables the kernels read AND the id -> row index live on the device: a hash table (csrc/store.cu)
answers get / remove by id in O(1) per id where the reference walks every list, and a removal is two kernels (erase +
per-list compaction that replays the reference's swap-with-last rule) with one small read-back of the new sizes.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import QkStore, check, ptr
from ._lib import QK_SEGMENT_ROWS as _MAX_SEGMENT_ROWS


def _round_up(x: int, a: int) -> int:
    return (x + a - 1) // a * a


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


_NEXT_UID = [0]


class PartitionStore:
    def __init__(self, d: int, device: torch.device):
        _NEXT_UID[0] += 1
        self.uid = _NEXT_UID[0]  # process-unique: (uid, version) identifies one state of one store
        self.d = int(d)
        self.pitch = _round_up(self.d, 4)
        self.device = device
        self.vectors = torch.zeros((0, self.pitch), dtype=torch.float32, device=device)
        self.ids = torch.zeros((0,), dtype=torch.int64, device=device)
        self.norms = torch.zeros((0,), dtype=torch.float32, device=device)  # squared row norms
        self.rows_used = 0
        self.list_row0 = np.zeros(0, dtype=np.int64)
        self.list_size = np.zeros(0, dtype=np.int64)
        self.list_cap = np.zeros(0, dtype=np.int64)
        self.slot_pid = np.zeros(0, dtype=np.int64)
        self.pid_slot: dict[int, int] = {}
        self.free_slots: list[int] = []
        self.curr_list_id = 0
        self.max_row_norm = 0.0
        # tensor-core filter precision of scans over this store (qk_store_t.filter_terms): 3 = 3xTF32, 2 = 2xTF32.
        # QuakeIndex sets 2 on the partition store of a two-level index (see index.py: filter precision policy).
        self.filter_terms = 3
        self._dirty = True
        self.version = 0  # bumped by every mutation; search plans (CUDA graphs) are keyed by it
        self._cache = {}
        # id -> arena row hash table on the device (built lazily; kept incrementally by append / remove, rebuilt after
        # anything that moves whole lists)
        self._hkeys = None
        self._hvals = None
        self._hash_valid = False
        self._hash_fill = 0  # live entries + tombstones

    # ------------------------------------------------------------------ basic queries
    @property
    def nlist(self) -> int:
        return len(self.pid_slot)

    @property
    def ntotal(self) -> int:
        return int(self.list_size.sum()) if self.list_size.size else 0

    def partition_ids(self) -> np.ndarray:
        return np.array(sorted(self.pid_slot.keys()), dtype=np.int64)

    def size_of(self, pid: int) -> int:
        return int(self.list_size[self.pid_slot[int(pid)]])

    # ------------------------------------------------------------------ arena management
    def _reserve_rows(self, rows: int) -> None:
        if rows <= self.vectors.shape[0]:
            return
        new_cap = max(rows, int(self.vectors.shape[0] * 1.5) + 1024)
        nv = torch.zeros((new_cap, self.pitch), dtype=torch.float32, device=self.device)
        ni = torch.full((new_cap,), -1, dtype=torch.int64, device=self.device)
        nn = torch.zeros((new_cap,), dtype=torch.float32, device=self.device)
        if self.rows_used:
            nv[: self.rows_used].copy_(self.vectors[: self.rows_used])
            ni[: self.rows_used].copy_(self.ids[: self.rows_used])
            nn[: self.rows_used].copy_(self.norms[: self.rows_used])
        self.vectors, self.ids, self.norms = nv, ni, nn
        self._dirty = True

    def _new_slot(self) -> int:
        if self.free_slots:
            return self.free_slots.pop()
        s = self.slot_pid.size
        for name in ("list_row0", "list_size", "list_cap"):
            setattr(self, name, np.append(getattr(self, name), 0))
        self.slot_pid = np.append(self.slot_pid, -1)
        return s

    def add_list(self, pid: int, capacity: int = 0) -> int:
        """DynamicInvertedLists::add_list: a new empty partition."""
        pid = int(pid)
        if pid in self.pid_slot:
            raise RuntimeError("List already exists in add_list")
        s = self._new_slot()
        cap = int(capacity)
        self.list_row0[s] = self.rows_used
        self.list_size[s] = 0
        self.list_cap[s] = cap
        self.slot_pid[s] = pid
        self.pid_slot[pid] = s
        self._reserve_rows(self.rows_used + cap)
        self.rows_used += cap
        self.curr_list_id = max(self.curr_list_id, pid + 1)
        self._dirty = True
        return s

    def remove_list(self, pid: int) -> None:
        pid = int(pid)
        s = self.pid_slot.pop(pid, None)
        if s is None:
            raise RuntimeError("List does not exist in remove_list")
        self.list_size[s] = 0
        self.list_cap[s] = 0  # the rows become dead space until compact()
        self.slot_pid[s] = -1
        self.free_slots.append(s)
        self._dirty = True
        self._invalidate_hash()

    def _grow_list(self, s: int, need: int) -> None:
        """Move list `s` to the end of the arena with capacity >= need (doubling)."""
        new_cap = max(need, 2 * int(self.list_cap[s]), 32)
        old0, n = int(self.list_row0[s]), int(self.list_size[s])
        self._reserve_rows(self.rows_used + new_cap)
        new0 = self.rows_used
        if n:
            self.vectors[new0:new0 + n].copy_(self.vectors[old0:old0 + n])
            self.ids[new0:new0 + n].copy_(self.ids[old0:old0 + n])
            self.norms[new0:new0 + n].copy_(self.norms[old0:old0 + n])
        self.list_row0[s] = new0
        self.list_cap[s] = new_cap
        self.rows_used += new_cap
        self._dirty = True
        self._invalidate_hash()  # rows moved

    # ------------------------------------------------------------------ bulk build
    def init_from_sorted(self, src: torch.Tensor, src_ids: torch.Tensor, order: torch.Tensor | None,
                         counts: np.ndarray, pids: np.ndarray, slack: bool = True) -> None:
        """Lay out len(counts) lists; list i receives rows src[order[offs[i]:offs[i+1]]]
        (PartitionManager::init_partitions, partition_manager.cpp:33-121)."""
        lib = _lib.load()
        counts = np.asarray(counts, dtype=np.int64)
        n = int(counts.sum())
        caps = counts + np.maximum(16, counts // 8) if slack else counts.copy()
        caps = (caps + 3) // 4 * 4
        row0 = np.concatenate([[0], np.cumsum(caps)[:-1]]).astype(np.int64) if len(caps) else np.zeros(0, np.int64)
        total = int(caps.sum())
        self.vectors = torch.zeros((total, self.pitch), dtype=torch.float32, device=self.device)
        self.ids = torch.full((total,), -1, dtype=torch.int64, device=self.device)
        self.norms = torch.zeros((total,), dtype=torch.float32, device=self.device)
        self.rows_used = total
        nl = len(counts)
        self.list_row0 = row0.copy()
        self.list_size = counts.copy()
        self.list_cap = caps.astype(np.int64)
        self.slot_pid = np.asarray(pids, dtype=np.int64).copy()
        self.pid_slot = {int(p): i for i, p in enumerate(self.slot_pid)}
        self.free_slots = []
        self.curr_list_id = int(self.slot_pid.max()) + 1 if nl else 0
        if n:
            offs = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
            shift = torch.from_numpy(row0 - offs[:-1]).to(self.device)
            dst_rows = torch.arange(n, dtype=torch.int64, device=self.device) + torch.repeat_interleave(
                shift, torch.from_numpy(counts).to(self.device))
            check(lib.qk_scatter_rows(ptr(src), src.stride(0), ptr(src_ids), ptr(order), ptr(dst_rows), n, self.d,
                                      ptr(self.vectors), self.pitch, ptr(self.ids), _stream()))
        self._dirty = True
        self._invalidate_hash()
        self.refresh_max_norm()

    def refresh_max_norm(self) -> None:
        lib = _lib.load()
        if self.rows_used == 0:
            self.max_row_norm = 0.0
            return
        out = torch.zeros(1, dtype=torch.float32, device=self.device)
        # dead rows are zero or stale copies of live rows: both are safe for an upper bound
        check(lib.qk_max_row_norm(ptr(self.vectors), self.rows_used, self.pitch, self.d, ptr(out), _stream()))
        self.max_row_norm = float(out.item())
        self._refresh_norms(0, self.rows_used)
        self._dirty = True

    def _refresh_norms(self, row0: int, n: int) -> None:
        """Recompute the squared norms of arena rows [row0, row0 + n)."""
        if n <= 0:
            return
        check(_lib.load().qk_row_sqnorms(ptr(self.vectors[row0:]), n, self.pitch, self.d, ptr(self.norms[row0:]),
                                         _stream()))

    # ------------------------------------------------------------------ append / remove
    def append(self, slots: torch.Tensor, x: torch.Tensor, x_ids: torch.Tensor) -> None:
        """Append x[i] to list slot slots[i], input order preserved inside every list
        (PartitionManager::add, partition_manager.cpp:245-258)."""
        lib = _lib.load()
        n = int(x.shape[0])
        if n == 0:
            return
        nslots = self.slot_pid.size
        slots32 = slots.to(torch.int32).contiguous()
        counts_d = torch.zeros(nslots, dtype=torch.int64, device=self.device)
        offsets_d = torch.zeros(nslots + 1, dtype=torch.int64, device=self.device)
        order_d = torch.empty(n, dtype=torch.int64, device=self.device)
        wsb = lib.qk_partition_workspace_bytes(n, nslots)
        ws = torch.empty(wsb, dtype=torch.uint8, device=self.device)
        check(lib.qk_partition_by_assignment(ptr(slots32), n, nslots, ptr(counts_d), ptr(offsets_d), ptr(order_d),
                                             ptr(ws), wsb, _stream()))
        counts = counts_d.cpu().numpy()
        if int(counts.sum()) != n:
            raise RuntimeError("List does not exist in add_entries")
        for s in np.nonzero(counts)[0]:
            need = int(self.list_size[s] + counts[s])
            if need > self.list_cap[s]:
                self._grow_list(int(s), need)
        base = torch.from_numpy(self.list_row0 + self.list_size).to(self.device) - offsets_d[:-1]
        dst_rows = torch.arange(n, dtype=torch.int64, device=self.device) + torch.repeat_interleave(base, counts_d)
        x = x.contiguous()
        check(lib.qk_scatter_rows(ptr(x), x.stride(0), ptr(x_ids.contiguous()), ptr(order_d), ptr(dst_rows), n, self.d,
                                  ptr(self.vectors), self.pitch, ptr(self.ids), _stream()))
        xn = torch.empty(n, dtype=torch.float32, device=self.device)
        check(lib.qk_row_sqnorms(ptr(x), n, x.stride(0), self.d, ptr(xn), _stream()))
        self.norms[dst_rows] = xn[order_d]
        self._hash_add(x_ids.to(torch.int64)[order_d], dst_rows)
        self.list_size = self.list_size + counts
        out = torch.tensor([self.max_row_norm], dtype=torch.float32, device=self.device)
        check(lib.qk_max_row_norm(ptr(x), n, x.stride(0), self.d, ptr(out), _stream()))
        self.max_row_norm = float(out.item())
        self._dirty = True

    def live_rows(self) -> torch.Tensor:
        """Arena rows of all live vectors, list slot by list slot."""
        sz = torch.from_numpy(self.list_size).to(self.device)
        r0 = torch.from_numpy(self.list_row0).to(self.device)
        n = int(self.list_size.sum())
        if n == 0:
            return torch.zeros(0, dtype=torch.int64, device=self.device)
        starts = torch.cumsum(sz, 0) - sz
        return torch.arange(n, dtype=torch.int64, device=self.device) + torch.repeat_interleave(r0 - starts, sz)

    # ------------------------------------------------------------------ id -> row index (device hash table)
    def _invalidate_hash(self) -> None:
        self._hash_valid = False

    def _ensure_hash(self) -> None:
        """(Re)build the id -> row table from the live rows: one kernel over ntotal ids."""
        if self._hash_valid:
            return
        lib = _lib.load()
        rows = self.live_rows()
        n = int(rows.numel())
        cap = int(lib.qk_hash_capacity(max(n, 1) * 2))  # room for as many inserts again before the next rebuild
        if self._hkeys is None or int(self._hkeys.numel()) != cap:
            self._hkeys = torch.empty(cap, dtype=torch.int64, device=self.device)
            self._hvals = torch.empty(cap, dtype=torch.int64, device=self.device)
        check(lib.qk_hash_clear(ptr(self._hkeys), cap, _stream()))
        if n:
            failed = torch.zeros(1, dtype=torch.int32, device=self.device)
            live_ids = self.ids[rows]
            check(lib.qk_hash_insert(ptr(self._hkeys), ptr(self._hvals), cap, ptr(live_ids), ptr(rows), n, ptr(failed), _stream()))
        self._hash_fill = n
        self._hash_valid = True

    def _hash_add(self, ids: torch.Tensor, rows: torch.Tensor) -> None:
        """Keep a valid table in step with an append / a moved list (insert or overwrite)."""
        if not self._hash_valid:
            return
        n = int(ids.numel())
        if self._hash_fill + n > int(self._hkeys.numel()) * 6 // 10:
            self._hash_valid = False  # too full: rebuilt (larger) on the next lookup
            return
        failed = torch.zeros(1, dtype=torch.int32, device=self.device)
        check(_lib.load().qk_hash_insert(ptr(self._hkeys), ptr(self._hvals), int(self._hkeys.numel()), ptr(ids.contiguous()),
                                         ptr(rows.contiguous()), n, ptr(failed), _stream()))
        self._hash_fill += n

    def find_rows(self, ids: torch.Tensor) -> torch.Tensor:
        """Arena row of each id (-1 if absent). The reference searches every list linearly
        (dynamic_inverted_list.cpp:302-321, index_partition.cpp:129-145); here one hash probe per id."""
        ids = ids.to(device=self.device, dtype=torch.int64).contiguous()
        out = torch.empty_like(ids)
        if ids.numel() == 0:
            return out
        self._ensure_hash()
        check(_lib.load().qk_hash_lookup(ptr(self._hkeys), ptr(self._hvals), int(self._hkeys.numel()), ptr(ids),
                                         int(ids.numel()), ptr(out), _stream()))
        return out

    def device_list_table(self):
        """(row0 [nslots], size [nslots]) int64 device copies of the host list table (cached per version)."""
        self.tables()
        if "list_table" not in self._cache:
            self._cache["list_table"] = (torch.from_numpy(self.list_row0.astype(np.int64)).to(self.device),
                                         torch.from_numpy(self.list_size.astype(np.int64)).to(self.device))
        return self._cache["list_table"]

    def device_sizes_by_pid(self) -> torch.Tensor:
        """int64 [curr_list_id] device tensor: size of partition id p (0 for ids that are not live)."""
        self.tables()
        if "sizes_by_pid" not in self._cache:
            t = np.zeros(max(self.curr_list_id, 1), dtype=np.int64)
            for pid, s in self.pid_slot.items():
                t[pid] = self.list_size[s]
            self._cache["sizes_by_pid"] = torch.from_numpy(t).to(self.device)
        return self._cache["sizes_by_pid"]

    def remove_ids(self, ids: torch.Tensor) -> int:
        """DynamicInvertedLists::remove_vectors (dynamic_inverted_list.cpp:137-149): every list drops its
        members found in `ids`, filling each hole with the list's current last element. Two kernels (erase from the
        id table + flag the rows; per-list compaction) and one read-back of the new list sizes; duplicate and absent
        ids are ignored like the reference's std::set lookup (partition_manager.cpp:306-310)."""
        lib = _lib.load()
        ids = ids.to(device=self.device, dtype=torch.int64).contiguous()
        n = int(ids.numel())
        nslots = self.slot_pid.size
        if n == 0 or nslots == 0 or self.rows_used == 0:
            return 0
        self._ensure_hash()
        rows_cap = int(self.vectors.shape[0])
        flags = torch.zeros(rows_cap, dtype=torch.uint8, device=self.device)
        n_erased = torch.zeros(1, dtype=torch.int64, device=self.device)
        check(lib.qk_store_remove(ptr(self._hkeys), ptr(self._hvals), int(self._hkeys.numel()), ptr(ids), n, None, ptr(flags),
                                  ptr(n_erased), _stream()))
        row0_d, size_d = self.device_list_table()
        new_size = torch.empty(nslots, dtype=torch.int64, device=self.device)
        scratch = torch.empty((2, rows_cap), dtype=torch.int32, device=self.device)
        check(lib.qk_store_compact_lists(ptr(self.vectors), self.pitch, ptr(self.ids), ptr(self.norms), ptr(row0_d), ptr(size_d),
                                         nslots, ptr(new_size), ptr(flags), ptr(scratch[0]), ptr(scratch[1]), ptr(self._hkeys),
                                         ptr(self._hvals), int(self._hkeys.numel()), _stream()))
        sizes = new_size.cpu().numpy()  # the one synchronisation of a removal
        live = self.slot_pid >= 0
        removed = int((self.list_size[live] - sizes[live]).sum())
        self.list_size = np.where(live, sizes, self.list_size).astype(np.int64)
        self._dirty = True
        return removed

    def all_ids(self) -> torch.Tensor:
        """Ids of all live vectors, partition by partition in ascending partition id (device; one gather)."""
        pids = self.partition_ids()
        if pids.size == 0:
            return torch.zeros(0, dtype=torch.int64, device=self.device)
        return self.ids[self.rows_of(pids)]

    def get_list(self, pid: int, padded: bool = False):
        s = self.pid_slot[int(pid)]
        r0, n = int(self.list_row0[s]), int(self.list_size[s])
        v = self.vectors[r0:r0 + n]
        return (v if padded else v[:, : self.d]), self.ids[r0:r0 + n]

    def set_list(self, pid: int, vecs: torch.Tensor, ids: torch.Tensor) -> None:
        """Replace the content of a list (kmeans_refine_partitions hands back rebuilt partitions)."""
        s = self.pid_slot[int(pid)]
        n = int(vecs.shape[0])
        if n > self.list_cap[s]:
            self.list_size[s] = 0
            self._grow_list(s, n + max(16, n // 8))
        r0 = int(self.list_row0[s])
        if n:
            self.vectors[r0:r0 + n, : self.d] = vecs
            if self.pitch > self.d:
                self.vectors[r0:r0 + n, self.d:] = 0
            self.ids[r0:r0 + n] = ids
            self._refresh_norms(r0, n)
            out = torch.tensor([self.max_row_norm], dtype=torch.float32, device=self.device)
            check(_lib.load().qk_max_row_norm(ptr(self.vectors[r0:]), n, self.pitch, self.d, ptr(out), _stream()))
            self.max_row_norm = float(out.item())
        self.list_size[s] = n
        self._dirty = True
        self._invalidate_hash()

    def rows_of(self, pids) -> torch.Tensor:
        """Arena rows of the members of the given partitions, partition by partition (device int64)."""
        slots = np.array([self.pid_slot[int(p)] for p in pids], dtype=np.int64)
        sz = torch.from_numpy(self.list_size[slots]).to(self.device)
        r0 = torch.from_numpy(self.list_row0[slots]).to(self.device)
        n = int(self.list_size[slots].sum())
        if n == 0:
            return torch.zeros(0, dtype=torch.int64, device=self.device)
        starts = torch.cumsum(sz, 0) - sz
        return torch.arange(n, dtype=torch.int64, device=self.device) + torch.repeat_interleave(r0 - starts, sz)

    def replace_lists(self, pids, counts: np.ndarray, vecs: torch.Tensor, ids: torch.Tensor) -> None:
        """Replace the content of several lists at once: list pids[j] receives rows
        vecs[offs[j]:offs[j+1]] (kmeans_refine_partitions hands back all rebuilt partitions together). The rows
        are a permutation of vectors already in the store, so the row-norm bound is unchanged."""
        slots = np.array([self.pid_slot[int(p)] for p in pids], dtype=np.int64)
        counts = np.asarray(counts, dtype=np.int64)
        self.list_size[slots] = 0
        for j in np.nonzero(counts > self.list_cap[slots])[0]:  # the few lists that outgrow their slack
            self._grow_list(int(slots[j]), int(counts[j] + max(16, counts[j] // 8)))
        n = int(counts.sum())
        if n:
            cnt_d = torch.from_numpy(counts).to(self.device)
            r0 = torch.from_numpy(self.list_row0[slots]).to(self.device)
            starts = torch.cumsum(cnt_d, 0) - cnt_d
            dst = torch.arange(n, dtype=torch.int64, device=self.device) + torch.repeat_interleave(r0 - starts, cnt_d)
            vecs = vecs.contiguous()
            check(_lib.load().qk_scatter_rows(ptr(vecs), vecs.stride(0), ptr(ids.contiguous()), None, ptr(dst), n, self.d,
                                              ptr(self.vectors), self.pitch, ptr(self.ids), _stream()))
            xn = torch.empty(n, dtype=torch.float32, device=self.device)
            check(_lib.load().qk_row_sqnorms(ptr(vecs), n, vecs.stride(0), self.d, ptr(xn), _stream()))
            self.norms[dst] = xn
        self.list_size[slots] = counts
        self._dirty = True
        self._invalidate_hash()

    def dead_rows(self) -> int:
        """Arena rows that belong to no live list (lists that moved or were removed leave their old runs behind)."""
        live = self.slot_pid >= 0
        return int(self.rows_used - self.list_cap[live].sum()) if self.slot_pid.size else 0

    def maybe_compact(self) -> bool:
        """Rewrite the arena once more than half of it is dead space."""
        if self.rows_used > 4096 and 2 * self.dead_rows() > self.rows_used:
            self.compact()
            return True
        return False

    def compact(self) -> None:
        """Rewrite the arena without dead space (lists keep their content order)."""
        pids = self.partition_ids()
        slots = np.array([self.pid_slot[int(p)] for p in pids], dtype=np.int64)
        counts = self.list_size[slots]
        order = self.rows_of(pids) if len(pids) else torch.zeros(0, dtype=torch.int64, device=self.device)
        old_v, old_i = self.vectors, self.ids
        norm, next_id, terms = self.max_row_norm, self.curr_list_id, self.filter_terms
        self.init_from_sorted(old_v, old_i, order, counts, pids)
        self.max_row_norm = max(norm, self.max_row_norm)
        self.curr_list_id = max(next_id, self.curr_list_id)  # partition ids are never reused
        self.filter_terms = terms

    # ------------------------------------------------------------------ device tables for the kernels
    def segment_len(self, num_queries: int, nprobe: int) -> int:
        """Scan-segment length for a batch. A work item of the scan kernel is (segment, chunk of <= 32 queries);
        IVF lists are short and numerous, so the default cut (QK_SEGMENT_ROWS) leaves them whole. A store of a
        few long lists (a flat index / the centroid list is ONE list) is cut finer until the batch yields a few
        items per SM."""
        import os
        if os.environ.get("QK_SEG_LEN"):
            return int(os.environ["QK_SEG_LEN"])
        if num_queries * nprobe < 1024:
            return 256
        if self.nlist > 16 or self.list_size.size == 0:
            return _MAX_SEGMENT_ROWS
        if os.environ.get("QK_FLAT_SEG_LEN"):  # experiments: few-list stores only (the coarse scan)
            return int(os.environ["QK_FLAT_SEG_LEN"])
        chunks = max(1, (num_queries + 31) // 32) * min(self.nlist, nprobe)
        longest = max(int(self.list_size.max()), 1)
        seg_len = _MAX_SEGMENT_ROWS
        while seg_len > 256 and chunks * ((longest + seg_len - 1) // seg_len) < 444:  # ~3 items per SM (coarse scan of C2: 30 us at 3.5 items per SM, 42 us at 1.7)
            seg_len //= 2
        return seg_len

    def set_filter_terms(self, terms: int) -> None:
        if int(terms) != self.filter_terms:
            self.filter_terms = int(terms)
            self._dirty = True  # the struct is rebuilt, captured plans are dropped (version bump)

    def tables_snapshot(self):
        """References to everything a kernel launch of the current version reads: keeps it alive for a captured graph."""
        return (self.vectors, self.ids, self.norms, dict(self._cache))

    def tables(self, seg_len: int = _MAX_SEGMENT_ROWS):
        """(QkStore struct, id_to_slot tensor) for scan segments of `seg_len` rows; rebuilt lazily after any
        mutation."""
        key = int(seg_len)
        if self._dirty:
            self._cache = {}
            self._dirty = False
            self.version += 1
        if key in self._cache:
            return self._cache[key]
        nslots = self.slot_pid.size
        size = self.list_size
        nseg = (size + seg_len - 1) // seg_len
        seg0 = np.concatenate([[0], np.cumsum(nseg)[:-1]]) if nslots else np.zeros(0, np.int64)
        S = int(nseg.sum())
        seg_list = np.repeat(np.arange(nslots), nseg)
        seg_idx = np.arange(S) - np.repeat(seg0, nseg)
        seg_row0 = self.list_row0[seg_list] + seg_idx * seg_len
        seg_rows = np.minimum(seg_len, size[seg_list] - seg_idx * seg_len)
        dev = self.device
        t = {
            "list_seg0": torch.from_numpy(seg0.astype(np.int32)).to(dev),
            "list_nseg": torch.from_numpy(nseg.astype(np.int32)).to(dev),
            "seg_row0": torch.from_numpy(seg_row0.astype(np.int64)).to(dev),
            "seg_rows": torch.from_numpy(seg_rows.astype(np.int32)).to(dev),
        }
        table_size = max(self.curr_list_id, 1)
        id_to_slot = np.full(table_size, -1, dtype=np.int32)
        for pid, s in self.pid_slot.items():
            id_to_slot[pid] = s
        t["id_to_slot"] = torch.from_numpy(id_to_slot).to(dev)
        st = QkStore()
        st.vectors = self.vectors.data_ptr()
        st.ids = self.ids.data_ptr()
        st.pitch = self.pitch
        st.d = self.d
        st.num_lists = nslots
        st.list_seg0 = t["list_seg0"].data_ptr()
        st.list_nseg = t["list_nseg"].data_ptr()
        st.num_segments = S
        st.max_list_segments = int(nseg.max()) if nslots else 0
        st.seg_row0 = t["seg_row0"].data_ptr()
        st.seg_rows = t["seg_rows"].data_ptr()
        st.max_row_norm = float(self.max_row_norm)
        st.row_norms = self.norms.data_ptr()
        st.num_rows = int(self.vectors.shape[0])
        st.flat_row0 = int(self.list_row0[0]) if nslots == 1 else 0
        st.flat_rows = int(self.list_size[0]) if nslots == 1 else 0
        st.max_segment_rows = int(seg_rows.max()) if S else 0
        st.filter_terms = int(self.filter_terms)
        st._keepalive = t  # the struct holds raw pointers into these tensors
        self._cache[key] = (st, t["id_to_slot"])
        return self._cache[key]
