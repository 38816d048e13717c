"""Parameter / result objects of the reference's Python surface.

Same attribute names and defaults as the pybind11 classes of the reference
(/root/reference/src/cpp/bindings/wrap.cpp:131-368; defaults /root/reference/src/cpp/include/common.h:70-99,
123-143, 171-247). Fields the reference binds but that have no meaning on a GPU (``num_workers``,
``num_threads``, ``aps_flush_period_us``) are accepted and ignored.
"""
from __future__ import annotations

import json

DEFAULT_NLIST = 0
DEFAULT_NITER = 5
DEFAULT_METRIC = "l2"
DEFAULT_NUM_WORKERS = 0
DEFAULT_K = 1
DEFAULT_NPROBE = 1
DEFAULT_RECALL_TARGET = -1.0
DEFAULT_BATCHED_SCAN = False
DEFAULT_PRECOMPUTED = True
DEFAULT_INITIAL_SEARCH_FRACTION = 0.02
DEFAULT_RECOMPUTE_THRESHOLD = 0.001
DEFAULT_APS_FLUSH_PERIOD_US = 100


def str_to_metric(metric: str) -> int:
    """common.h:145-156 -- 'l2' / 'ip' (case-insensitive), anything else is std::invalid_argument."""
    m = str(metric).lower()
    if m == "l2":
        return 1  # faiss::METRIC_L2
    if m == "ip":
        return 0  # faiss::METRIC_INNER_PRODUCT
    raise ValueError("Invalid metric type: " + str(metric))


def metric_to_str(metric: int) -> str:
    if metric == 1:
        return "l2"
    if metric == 0:
        return "ip"
    raise ValueError("Invalid metric type")


class _Repr:
    _repr_fields: tuple = ()

    def __repr__(self):
        return json.dumps({f: getattr(self, f) for f in self._repr_fields})


class IndexBuildParams(_Repr):
    _repr_fields = ("nlist", "niter", "metric", "num_workers")

    def __init__(self):
        self.nlist = DEFAULT_NLIST
        self.niter = DEFAULT_NITER
        self.metric = DEFAULT_METRIC
        self.num_workers = DEFAULT_NUM_WORKERS


class SearchParams(_Repr):
    _repr_fields = ("k", "nprobe", "recall_target", "batched_scan", "use_precomputed", "initial_search_fraction",
                    "recompute_threshold", "aps_flush_period_us")

    def __init__(self):
        self.k = DEFAULT_K
        self.nprobe = DEFAULT_NPROBE
        self.recall_target = DEFAULT_RECALL_TARGET
        self.num_threads = 1
        self.batched_scan = DEFAULT_BATCHED_SCAN
        self.use_precomputed = DEFAULT_PRECOMPUTED
        self.initial_search_fraction = DEFAULT_INITIAL_SEARCH_FRACTION
        self.recompute_threshold = DEFAULT_RECOMPUTE_THRESHOLD
        self.aps_flush_period_us = DEFAULT_APS_FLUSH_PERIOD_US


class MaintenancePolicyParams(_Repr):
    _repr_fields = ("maintenance_policy", "window_size", "refinement_radius", "refinement_iterations",
                    "min_partition_size", "alpha", "enable_split_rejection", "enable_delete_rejection",
                    "delete_threshold_ns", "split_threshold_ns")

    def __init__(self):
        self.maintenance_policy = "query_cost"
        self.window_size = 1000
        self.refinement_radius = 25
        self.refinement_iterations = 3
        self.min_partition_size = 32
        self.alpha = 0.9
        self.enable_split_rejection = True
        self.enable_delete_rejection = True
        self.delete_threshold_ns = 10.0
        self.split_threshold_ns = 10.0


class SearchTimingInfo(_Repr):
    _repr_fields = ("total_time_ns", "buffer_init_time_ns", "job_enqueue_time_ns", "boundary_distance_time_ns",
                    "job_wait_time_ns", "result_aggregate_time_ns", "n_queries", "n_clusters", "partitions_scanned")

    def __init__(self):
        self.total_time_ns = 0
        self.buffer_init_time_ns = 0
        self.job_enqueue_time_ns = 0
        self.boundary_distance_time_ns = 0
        self.job_wait_time_ns = 0
        self.result_aggregate_time_ns = 0
        self.n_queries = 0
        self.n_clusters = 0
        self.partitions_scanned = 0
        self.search_params = None
        self.parent_info = None


class SearchResult:
    def __init__(self):
        self.ids = None
        self.distances = None
        self.timing_info = None

    def __repr__(self):
        ni = 0 if self.ids is None else self.ids.numel()
        nd = 0 if self.distances is None else self.distances.numel()
        return json.dumps({"num_ids": ni, "num_distances": nd})


class BuildTimingInfo(_Repr):
    _repr_fields = ("total_time_us", "assign_time_us", "train_time_us", "d", "code_size", "n_codebooks", "n_vectors")

    def __init__(self):
        self.total_time_us = 0
        self.assign_time_us = 0
        self.train_time_us = 0
        self.d = 0
        self.code_size = 0
        self.n_codebooks = 0
        self.n_vectors = 0
        self.n_clusters = 0


class ModifyTimingInfo(_Repr):
    _repr_fields = ("modify_count", "input_validation_time_us", "modify_time_us", "find_partition_time_us")

    def __init__(self):
        self.modify_count = 0
        self.input_validation_time_us = 0
        self.modify_time_us = 0
        self.find_partition_time_us = 0


class MaintenanceTimingInfo(_Repr):
    _repr_fields = ("total_time_us", "split_time_us", "delete_time_us", "split_refine_time_us",
                    "delete_refine_time_us", "n_splits", "n_deletes")

    def __init__(self):
        self.total_time_us = 0
        self.split_time_us = 0
        self.delete_time_us = 0
        self.split_refine_time_us = 0
        self.delete_refine_time_us = 0
        self.n_splits = 0
        self.n_deletes = 0
