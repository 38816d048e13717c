"""ctypes binding of the C ABI declared in include/quake_b200.h.

The shared library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a) as
``quake_b200/lib/libquake_b200.so``. There is no CPU fallback: if the library is missing, or a compute
entry point is called without a B200, the call raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QK_LIB_PATH") or os.path.join(_HERE, "lib", "libquake_b200.so")  # QK_LIB_PATH: profiling builds

QK_METRIC_INNER_PRODUCT = 0
QK_METRIC_L2 = 1
QK_SEGMENT_ROWS = 4096
QK_MAX_K = 2048

c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_f32p = C.POINTER(C.c_float)
vp = C.c_void_p


class QkStore(C.Structure):
    """struct qk_store (include/quake_b200.h)."""

    _fields_ = [
        ("vectors", vp),
        ("ids", vp),
        ("pitch", C.c_int64),
        ("d", C.c_int32),
        ("num_lists", C.c_int32),
        ("list_seg0", vp),
        ("list_nseg", vp),
        ("num_segments", C.c_int32),
        ("max_list_segments", C.c_int32),
        ("seg_row0", vp),
        ("seg_rows", vp),
        ("max_row_norm", C.c_float),
        ("row_norms", vp),
        ("num_rows", C.c_int64),
        ("flat_row0", C.c_int64),
        ("flat_rows", C.c_int64),
        ("max_segment_rows", C.c_int32),
        ("filter_terms", C.c_int32),
    ]


# name -> (restype, argtypes); every symbol include/quake_b200.h declares
PROTOTYPES = {
    "qk_version": (C.c_char_p, []),
    "qk_last_error": (C.c_char_p, []),
    "qk_launch_count": (C.c_longlong, []),
    "qk_device_check": (C.c_int, [c_i32p, c_i32p, c_i32p]),
    "qk_scan_workspace_bytes": (C.c_size_t, [C.POINTER(QkStore), C.c_int64, C.c_int, C.c_int]),
    "qk_scan_partitions": (
        C.c_int,
        [C.POINTER(QkStore), vp, C.c_int64, C.c_int64, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, C.c_size_t, vp, vp],
    ),
    "qk_search_ivf_workspace_bytes": (C.c_size_t, [C.POINTER(QkStore), C.POINTER(QkStore), C.c_int64, C.c_int, C.c_int]),
    "qk_search_ivf": (
        C.c_int,
        [C.POINTER(QkStore), C.POINTER(QkStore), vp, C.c_int64, vp, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int,
         C.c_int, vp, vp, vp, vp, C.c_size_t, vp, vp],
    ),
    "qk_profile_begin": (C.c_int, [C.c_int]),
    "qk_profile_count": (C.c_int, []),
    "qk_profile_read": (C.c_int, [C.c_int, c_f32p, c_i64p, c_i32p, c_i32p]),
    "qk_profile_end": (C.c_int, []),
    "qk_map_ids_to_slots": (C.c_int, [vp, C.c_int64, vp, C.c_int64, vp, vp]),
    "qk_max_row_norm": (C.c_int, [vp, C.c_int64, C.c_int64, C.c_int, vp, vp]),
    "qk_row_sqnorms": (C.c_int, [vp, C.c_int64, C.c_int64, C.c_int, vp, vp]),
    "qk_merge_topk": (C.c_int, [vp, vp, C.c_int, C.c_int64, C.c_int, C.c_int, vp, vp, vp]),
    "qk_peer_buffer_bytes": (C.c_size_t, [C.c_int64, C.c_int, C.c_int]),
    "qk_peer_alloc": (C.c_int, [C.c_size_t, C.POINTER(vp), vp]),
    "qk_peer_open": (C.c_int, [vp, C.POINTER(vp)]),
    "qk_peer_close": (C.c_int, [vp]),
    "qk_peer_free": (C.c_int, [vp]),
    "qk_exchange_merge_topk": (C.c_int, [vp, vp, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp), vp, vp, vp]),
    "qk_host_beta_table": (C.c_int, [C.c_int, C.POINTER(C.c_double)]),
    "qk_aps_boundary_distances": (C.c_int, [vp, C.c_int64, C.c_int64, C.c_int, vp, C.c_int64, vp, C.c_int, C.c_int, vp, vp]),
    "qk_aps_thresholds": (C.c_int, [vp, C.c_int64, vp, C.c_int64, C.c_int, vp, vp, C.c_int, C.c_int, C.c_float, C.c_int, vp, vp]),
    "qk_scan_collect": (
        C.c_int,
        [C.POINTER(QkStore), vp, C.c_int64, C.c_int64, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, C.c_size_t, vp],
    ),
    "qk_aps_advance": (
        C.c_int,
        [vp, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, C.c_float, C.c_float,
         C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp],
    ),
    "qk_kmeans_assign_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int64, C.c_int]),
    "qk_kmeans_assign": (
        C.c_int,
        [vp, C.c_int64, C.c_int64, C.c_int, vp, C.c_int64, C.c_int64, C.c_int, vp, vp, vp, C.c_size_t, vp],
    ),
    "qk_kmeans_assign_filtered": (
        C.c_int,
        [vp, C.c_int64, C.c_int64, C.c_int, vp, C.c_int64, C.c_int64, C.c_int, C.c_int, vp, vp, vp, vp, C.c_size_t, vp],
    ),
    "qk_kmeans_accumulate": (C.c_int, [vp, C.c_int64, C.c_int, vp, vp, C.c_int64, vp, C.c_int64, vp]),
    "qk_partition_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int64]),
    "qk_partition_by_assignment": (C.c_int, [vp, C.c_int64, C.c_int64, vp, vp, vp, vp, C.c_size_t, vp]),
    "qk_gather_rows": (C.c_int, [vp, C.c_int64, vp, vp, C.c_int64, C.c_int, vp, C.c_int64, vp, vp]),
    "qk_scatter_rows": (C.c_int, [vp, C.c_int64, vp, vp, vp, C.c_int64, C.c_int, vp, C.c_int64, vp, vp]),
    "qk_normalize_rows": (C.c_int, [vp, C.c_int64, C.c_int64, C.c_int, vp]),
    "qk_hash_capacity": (C.c_int64, [C.c_int64]),
    "qk_hash_clear": (C.c_int, [vp, C.c_int64, vp]),
    "qk_hash_insert": (C.c_int, [vp, vp, C.c_int64, vp, vp, C.c_int64, vp, vp]),
    "qk_hash_lookup": (C.c_int, [vp, vp, C.c_int64, vp, C.c_int64, vp, vp]),
    "qk_store_remove": (C.c_int, [vp, vp, C.c_int64, vp, C.c_int64, vp, vp, vp, vp]),
    "qk_store_compact_lists": (C.c_int, [vp, C.c_int64, vp, vp, vp, vp, C.c_int64, vp, vp, vp, vp, vp, vp, C.c_int64, vp]),
    "qk_host_rand_perm_prefix": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, c_i64p]),
    "qk_host_split_clusters": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, c_f32p, c_f32p, C.c_int64, c_i64p]),
}

_lib = None


class QuakeB200Error(RuntimeError):
    pass


def load():
    """Load libquake_b200.so (once) and attach the prototypes. Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise QuakeB200Error(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "quake_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    """Translate a non-zero status into the Python exception class the reference would raise
    (std::invalid_argument -> ValueError, std::runtime_error -> RuntimeError; wrap.cpp / pybind11)."""
    if rc == 0:
        return
    msg = load().qk_last_error().decode("utf-8", "replace")
    if rc == 1:
        raise ValueError(msg)
    raise QuakeB200Error(f"[quake_b200 rc={rc}] {msg}")


_device_ok = False


def require_device() -> None:
    """Fail loudly unless a compute-capability-10.x GPU is current."""
    global _device_ok
    if _device_ok:
        return
    lib = load()
    sm, maj, mnr = C.c_int32(), C.c_int32(), C.c_int32()
    check(lib.qk_device_check(C.byref(sm), C.byref(maj), C.byref(mnr)))
    _device_ok = True


def ptr(t):
    """Device/host pointer of a torch tensor (or None)."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())
