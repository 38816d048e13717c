"""CPU-only tests of host-side logic that needs no device: the scan-segment tables the kernels read
(quake_b200/store.py), the segment-length heuristic, and the host half of APS (qk_host_beta_table) against the
oracle's restatement of geometry.h."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import oracle as orc
from quake_b200 import _lib
from quake_b200.store import PartitionStore


def _host_store(sizes, d=8, slack=4):
    """A PartitionStore with host tensors and hand-laid lists (no kernel involved)."""
    st = PartitionStore(d, torch.device("cpu"))
    sizes = np.asarray(sizes, dtype=np.int64)
    caps = sizes + slack
    row0 = np.concatenate([[0], np.cumsum(caps)[:-1]]).astype(np.int64)
    total = int(caps.sum())
    st.vectors = torch.zeros((total, st.pitch))
    st.ids = torch.full((total,), -1, dtype=torch.int64)
    st.norms = torch.zeros(total)
    st.rows_used = total
    st.list_row0, st.list_size, st.list_cap = row0, sizes.copy(), caps
    st.slot_pid = np.arange(len(sizes), dtype=np.int64) * 3 + 1  # partition ids 1, 4, 7, ...
    st.pid_slot = {int(p): i for i, p in enumerate(st.slot_pid)}
    st.curr_list_id = int(st.slot_pid.max()) + 1
    st._dirty = True
    return st


@pytest.mark.parametrize("seg_len", [256, 1024, 4096])
def test_segment_tables_cover_every_list_exactly(seg_len):
    sizes = [0, 1, 255, 256, 257, 5000, 4096, 4097, 9000]
    st = _host_store(sizes)
    q, table = st.tables(seg_len)
    t = q._keepalive
    seg0, nseg = t["list_seg0"].numpy(), t["list_nseg"].numpy()
    srow0, srows = t["seg_row0"].numpy(), t["seg_rows"].numpy()
    assert q.num_lists == len(sizes) and q.num_segments == int(nseg.sum()) == len(srows)
    assert q.max_list_segments == int(nseg.max()) and q.max_segment_rows == int(srows.max())
    for l, n in enumerate(sizes):
        assert nseg[l] == (n + seg_len - 1) // seg_len
        rows = []
        for s in range(seg0[l], seg0[l] + nseg[l]):
            assert 1 <= srows[s] <= seg_len
            rows += list(range(srow0[s], srow0[s] + srows[s]))
        assert rows == list(range(st.list_row0[l], st.list_row0[l] + n))  # contiguous, in order, nothing else
    # partition id -> slot table
    tab = table.numpy()
    assert tab.shape[0] == st.curr_list_id
    for pid, s in st.pid_slot.items():
        assert tab[pid] == s
    assert (tab >= 0).sum() == len(sizes)
    assert (q.flat_row0, q.flat_rows) == (0, 0)  # more than one list


def test_flat_store_geometry_and_version():
    st = _host_store([5000])
    q, _ = st.tables(512)
    assert (q.flat_row0, q.flat_rows, q.num_segments) == (0, 5000, 10)
    v = st.version
    st.tables(512)
    assert st.version == v  # cached
    st._dirty = True
    st.tables(512)
    assert st.version == v + 1  # a mutation invalidates search plans keyed by the version


def test_segment_len_heuristic():
    ivf = _host_store([244] * 64)
    assert ivf.segment_len(1024, 64) == 4096          # many short lists: leave them whole
    assert ivf.segment_len(4, 8) == 256               # tiny batch: spread the work
    coarse = _host_store([4096])
    assert coarse.segment_len(1024, 1) == 256         # the C2 coarse scan: 32 chunks x 16 segments (~3.5 items per SM)
    big = _host_store([65536])
    assert big.segment_len(1024, 1) >= 2048           # already plenty of items
    assert coarse.segment_len(32768, 1) == 4096       # k-means assign batches: chunks alone fill the GPU


@pytest.mark.parametrize("d", [16, 96, 128])
def test_host_beta_table_matches_oracle(d):
    """qk_host_beta_table = incomplete_beta_lookup's table (geometry.h:163-186): I_x((d+1)/2, 1/2) at x = i/1000,
    against the oracle's restatement (itself pinned to the compiled reference in test_oracle.py)."""
    lib = _lib.load()
    arr = np.empty(1001, dtype=np.float64)
    _lib.check(lib.qk_host_beta_table(d, arr.ctypes.data_as(C.POINTER(C.c_double))))
    want = np.array([orc.incomplete_beta((d + 1) / 2.0, 0.5, i / 1000.0) for i in range(1001)])
    # double precision, same formulas; the last bits depend on the host compiler's FMA contraction (the compiled
    # reference, the oracle and this library differ from one another by <= 6e-14 relative), far below what can move
    # the float probabilities APS compares
    assert np.allclose(arr, want, rtol=1e-12, atol=0)
    assert arr[0] == 0.0 and abs(arr[-1] - 1.0) < 1e-12 and np.all(np.diff(arr) >= 0)
