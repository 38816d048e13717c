"""Dynamic workload generation and replay against the GPU index.

Mirror of the reference's regression / ablation harness (/root/reference/src/python/workload_generator.py:42-606,
src/python/index_wrappers/quake.py:11-260, src/python/utils.py:162-240) so that its experiments run unchanged on this
index: the same classes (``DynamicWorkloadGenerator``, ``WorkloadEvaluator``, ``QuakeWrapper``, the two samplers),
constructor arguments and result dictionaries, and the same on-disk workload layout --

    <workload_dir>/runbook.json            parameters, per-operation entries {type, sample_size, n_resident, ...}, summary
    <workload_dir>/operations/<i>.pt       ids of the vectors / queries of operation i
    <workload_dir>/operations/<i>_gt_ids.pt, <i>_gt_dists.pt   exact top-100 over the resident set at that moment
    <workload_dir>/initial_indices.pt, base_vectors.pt, query_vectors.pt

-- so a workload written by either generator can be replayed by either evaluator. The random draws are made in the
reference's order (np.random.choice for the operation type, torch.randperm inside the samplers). Differences: the
ground truth is computed on the GPU (chunked matmul), plots are optional (matplotlib is not a dependency), and the
evaluator reports device-synchronised latencies.
"""
from __future__ import annotations

import json
import time
from pathlib import Path
from typing import Optional, Union

import numpy as np
import torch

from .index import QuakeIndex
from .params import IndexBuildParams, SearchParams


# ------------------------------------------------------------------------------------------------ utils.py
def to_path(p: Union[str, Path]) -> Path:
    return p if isinstance(p, Path) else Path(p)


def to_torch(x) -> torch.Tensor:
    return x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x))


def compute_recall(ids: torch.Tensor, gt_ids: torch.Tensor, k: int) -> torch.Tensor:
    """Per-query recall@k (utils.py:162-177)."""
    ids, gt_ids = to_torch(ids)[:, :k].cpu(), to_torch(gt_ids)[:, :k].cpu()
    assert ids.size() == gt_ids.size(), (ids.shape, gt_ids.shape)
    out = torch.zeros(ids.size(0))
    for i in range(ids.size(0)):
        out[i] = len(set(ids[i].tolist()) & set(gt_ids[i].tolist())) / k
    return out


def knn(queries, vectors, k: int = 1, metric: str = "l2", chunk: int = 1 << 20):
    """Exact k nearest neighbours (ids, distances) by brute force (utils.py:194-240); k = -1 ranks everything. Runs
    on the GPU when one is there (this is the harness' ground truth, not the index under test)."""
    queries, vectors = to_torch(queries).to(torch.float32), to_torch(vectors).to(torch.float32)
    if queries.dim() == 1:
        queries = queries.unsqueeze(0)
    assert vectors.dim() == 2 and queries.size(1) == vectors.size(1)
    dev = torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")
    n = vectors.size(0)
    k = n if k < 0 else min(k, n)
    q = queries.to(dev)
    qn = (q * q).sum(1, keepdim=True)
    best_d = best_i = None
    for s in range(0, n, chunk):
        v = vectors[s:s + chunk].to(dev)
        ip = q @ v.T
        score = -ip if metric == "ip" else (qn + (v * v).sum(1)[None, :] - 2 * ip)
        d, i = score.topk(min(k, v.size(0)), largest=False)
        i = i + s
        if best_d is not None:
            d, i = torch.cat([best_d, d], 1), torch.cat([best_i, i], 1)
            sel = d.topk(min(k, d.size(1)), largest=False).indices
            d, i = d.gather(1, sel), i.gather(1, sel)
        best_d, best_i = d, i
    dist = -best_d if metric == "ip" else best_d.clamp_min(0).sqrt()
    return best_i.cpu(), dist.cpu()


# ------------------------------------------------------------------------------------------------ index wrapper
class QuakeWrapper:
    """index_wrappers/quake.py: the thin keyword-argument facade the harness drives."""

    def __init__(self):
        self.index: QuakeIndex | None = None

    def n_total(self) -> int:
        return self.index.ntotal()

    def d(self) -> int:
        return self.index.d()

    def index_state(self) -> dict:
        return {"n_list": self.index.nlist(), "n_total": self.index.ntotal()}

    def build(self, vectors, nc: int, metric: str = "l2", ids=None, num_workers: int = 0, m: int = -1, code_size: int = 8):
        vectors = to_torch(vectors)
        assert vectors.ndim == 2 and nc > 0
        bp = IndexBuildParams()
        bp.metric, bp.nlist, bp.num_workers = metric.lower(), nc, num_workers
        self.index = QuakeIndex()
        if ids is None:
            ids = torch.arange(vectors.shape[0], dtype=torch.int64)
        return self.index.build(vectors, to_torch(ids).to(torch.int64), bp)

    def add(self, vectors, ids=None, num_threads: int = 0):
        vectors = to_torch(vectors)
        assert self.index is not None and vectors.ndim == 2
        if ids is None:
            cur = self.n_total()
            ids = torch.arange(cur, cur + vectors.shape[0], dtype=torch.int64)
        return self.index.add(vectors, to_torch(ids).to(torch.int64))

    def remove(self, ids):
        ids = to_torch(ids)
        assert self.index is not None and ids.ndim == 1
        return self.index.remove(ids.to(torch.int64))

    def search(self, query, k: int, nprobe: int = 1, batched_scan=False, recall_target: float = -1, k_factor=4.0,
               use_precomputed=True, initial_search_fraction=0.05, recompute_threshold=0.1, aps_flush_period_us=50,
               n_threads=1):
        sp = SearchParams()
        sp.nprobe, sp.recall_target, sp.use_precomputed, sp.batched_scan = nprobe, recall_target, use_precomputed, batched_scan
        sp.initial_search_fraction, sp.recompute_threshold = initial_search_fraction, recompute_threshold
        sp.aps_flush_period_us, sp.k, sp.num_threads = aps_flush_period_us, k, n_threads
        return self.index.search(to_torch(query), sp)

    def maintenance(self):
        return self.index.maintenance()

    def save(self, filename):
        self.index.save(str(filename))

    def load(self, filename, n_workers: int = 0, use_numa: bool = False, verbose: bool = False, verify_numa: bool = False,
             same_core: bool = True, use_centroid_workers: bool = False, use_adaptive_n_probe: bool = False):
        """index_wrappers/quake.py:166-187. The worker / NUMA switches of the reference's CPU engine are accepted and
        have no meaning here (as in the reference's own wrapper, which only prints them)."""
        self.index = QuakeIndex()
        self.index.load(str(filename), n_workers)

    def centroids(self) -> torch.Tensor:
        return self.index.parent.get(self.index.parent.get_ids())

    def cluster_ids(self) -> torch.Tensor:
        """index_wrappers/quake.py:198-204 forwards to ``index.cluster_assignments()``, which the reference's bindings
        (wrap.cpp:57-186) do not define; kept for the surface, with the same outcome."""
        return self.index.cluster_assignments()

    def metric(self) -> str:
        return "ip" if self.index.metric == 0 else "l2"


# ------------------------------------------------------------------------------------------------ samplers
class UniformSampler:
    def sample(self, sample_pool: torch.Tensor, size: int, update_ranks: bool = True):
        return sample_pool[torch.randperm(sample_pool.shape[0])[:size]]


class StratifiedClusterSampler:
    """Cluster after cluster, nearest-first from a moving root cluster (workload_generator.py:62-124): the skewed
    insert / delete / query streams of the reference's dynamic experiments."""

    def __init__(self, assignments: torch.Tensor, centroids: torch.Tensor):
        self.assignments, self.centroids = assignments, centroids
        present = torch.unique(assignments)
        self.update_ranks(present[torch.randint(0, present.shape[0], (1,))])

    def update_ranks(self, root_cluster) -> None:
        self.root_cluster = root_cluster
        ids, _ = knn(self.centroids[int(root_cluster)], self.centroids, -1, "l2")
        self.cluster_ranks = ids.flatten()

    def sample(self, sample_pool: torch.Tensor, size: int, update_ranks: bool = True):
        pool_assign = self.assignments[sample_pool]
        present = set(pool_assign.tolist())
        order = [c for c in self.cluster_ranks.tolist() if c in present]
        picked, have = [], 0
        for c in order:
            members = (pool_assign == c).nonzero(as_tuple=True)[0]
            if members.numel() == 0:
                continue
            take = min(size - have, members.numel())
            picked.append(sample_pool[members[torch.randperm(members.numel())[:take]]])
            have += take
            if have >= size:
                break
        out = torch.cat(picked) if picked else torch.tensor([], dtype=torch.long)
        if update_ranks and len(order) > 1:
            self.update_ranks(order[1])
        return torch.unique(out)


# ------------------------------------------------------------------------------------------------ generator
class DynamicWorkloadGenerator:
    """workload_generator.py:127-398: cluster the base vectors, pick an initial resident set, then draw a stream of
    insert / delete / query operations and save each with its ground truth."""

    def __init__(self, workload_dir, base_vectors, metric: str, insert_ratio: float, delete_ratio: float,
                 query_ratio: float, update_batch_size: int, query_batch_size: int, number_of_operations: int,
                 initial_size: int, cluster_size: int, cluster_sample_distribution: str, queries,
                 query_cluster_sample_distribution: str = "uniform", seed: int = 1738,
                 initial_clustering_path=None, overwrite: bool = False):
        self.workload_dir = to_path(workload_dir)
        self.base_vectors = to_torch(base_vectors).to(torch.float32)
        self.metric = metric.lower()
        self.insert_ratio, self.delete_ratio, self.query_ratio = insert_ratio, delete_ratio, query_ratio
        self.update_batch_size, self.query_batch_size = update_batch_size, query_batch_size
        self.number_of_operations, self.initial_size, self.cluster_size = number_of_operations, initial_size, cluster_size
        self.cluster_sample_distribution = cluster_sample_distribution
        self.query_cluster_sample_distribution = query_cluster_sample_distribution
        self.queries = None if queries is None else to_torch(queries).to(torch.float32)
        self.seed = seed
        self.initial_clustering_path = to_path(initial_clustering_path) if initial_clustering_path else None
        torch.manual_seed(seed)
        np.random.seed(seed)
        self.validate_parameters()
        self.workload_dir.mkdir(parents=True, exist_ok=True)
        self.operations_dir = self.workload_dir / "operations"
        self.operations_dir.mkdir(parents=True, exist_ok=True)
        n = self.base_vectors.shape[0]
        self.resident_set = torch.zeros(n, dtype=torch.bool)
        self.all_ids = torch.arange(n)
        self.assignments = None
        self.runbook = {}
        self.clustered_index = None
        self.sampler = self.query_sampler = None
        self.resident_history = []

    def workload_exists(self) -> bool:
        return (self.workload_dir / "runbook.json").exists()

    def validate_parameters(self) -> None:
        assert self.metric in ["l2", "ip"]
        for r in (self.insert_ratio, self.delete_ratio, self.query_ratio):
            assert 0 <= r <= 1
        assert abs(self.insert_ratio + self.delete_ratio + self.query_ratio - 1) < 1e-9
        assert self.update_batch_size > 0 and self.query_batch_size > 0 and self.number_of_operations > 0
        assert self.initial_size > 0 and self.cluster_size > 0
        assert self.cluster_sample_distribution in ["uniform", "skewed", "skewed_fixed"]

    def initialize_clustered_index(self) -> QuakeWrapper:
        index_dir = self.initial_clustering_path or (self.workload_dir / "clustered_index.bin")
        index = QuakeWrapper()
        if index_dir.exists():
            index.load(index_dir)
        else:
            n = self.base_vectors.shape[0]
            index.build(self.base_vectors, nc=max(n // self.cluster_size, 1), metric=self.metric, ids=torch.arange(n))
            index.save(str(self.workload_dir / "clustered_index.bin"))
        sp = SearchParams()
        sp.k, sp.batched_scan = 1, True
        self.assignments = index.index.parent.search(self.base_vectors, sp).ids.flatten()
        return index

    def sample(self, size: int, operation_type: str):
        if operation_type == "insert":
            pool = self.all_ids[~self.resident_set]
        elif operation_type == "delete":
            pool = self.all_ids[self.resident_set]
        elif operation_type == "query":
            pool = torch.arange(self.queries.shape[0]) if self.queries is not None else self.all_ids[~self.resident_set]
        else:
            raise ValueError(f"Invalid operation type {operation_type}.")
        if pool.shape[0] == 0:
            return torch.tensor([], dtype=torch.long)
        if operation_type in ("insert", "delete"):
            return self.sampler.sample(pool, size)
        return self.query_sampler.sample(pool, size, update_ranks=True)

    def initialize_workload(self) -> None:
        cents = self.clustered_index.centroids()
        if self.cluster_sample_distribution in ("skewed", "skewed_fixed"):
            self.sampler = StratifiedClusterSampler(self.assignments, cents)
        else:
            self.sampler = UniformSampler()
        if self.query_cluster_sample_distribution in ("skewed", "skewed_fixed"):
            q_assign = knn(self.queries, cents, 1, "l2")[0].flatten()
            self.query_sampler = StratifiedClusterSampler(q_assign, cents)
        elif self.query_cluster_sample_distribution == "uniform":
            self.query_sampler = UniformSampler()
        else:
            raise ValueError(f"Invalid query cluster sample distribution {self.query_cluster_sample_distribution}.")
        initial = self.sample(self.initial_size, "insert")
        self.resident_set[initial] = True
        torch.save(initial, self.workload_dir / "initial_indices.pt")
        if self.queries is not None:
            torch.save(self.queries, self.workload_dir / "query_vectors.pt")
        torch.save(self.base_vectors, self.workload_dir / "base_vectors.pt")
        self.runbook["parameters"] = {
            "sample_queries": self.queries is None, "n_base_vectors": self.base_vectors.shape[0],
            "vector_dimension": self.base_vectors.shape[1], "metric": self.metric, "insert_ratio": self.insert_ratio,
            "delete_ratio": self.delete_ratio, "query_ratio": self.query_ratio,
            "update_batch_size": self.update_batch_size, "query_batch_size": self.query_batch_size,
            "number_of_operations": self.number_of_operations, "initial_size": self.initial_size,
            "cluster_size": self.cluster_size, "cluster_sample_distribution": self.cluster_sample_distribution,
            "query_cluster_sample_distribution": self.query_cluster_sample_distribution, "seed": self.seed}
        self.runbook["initialize"] = {"size": self.initial_size}
        self.runbook["operations"] = {}

    def generate_workload(self) -> dict:
        self.clustered_index = self.initialize_clustered_index()
        self.initialize_workload()
        counts = {"insert": 0, "delete": 0, "query": 0}
        n_operations = 0
        uniq, cnt = torch.unique(self.assignments, return_counts=True)
        all_sizes = torch.zeros(int(self.assignments.max().item()) + 1)
        all_sizes[uniq] = cnt.float()
        for i in range(self.number_of_operations):
            op = str(np.random.choice(["insert", "delete", "query"],
                                      p=[self.insert_ratio, self.delete_ratio, self.query_ratio]))
            counts[op] += 1
            ids = self.sample(self.query_batch_size if op == "query" else self.update_batch_size, op)
            if ids.shape[0] == 0:
                break
            n_operations = i + 1
            if op in ("insert", "delete"):
                self.resident_set[ids] = (op == "insert")
            n_resident = int(self.resident_set.sum().item())
            if n_resident < 5 * self.update_batch_size:
                print(f"Below minimum resident set size: {n_resident}")
                break
            entry = {"type": op, "sample_size": int(ids.shape[0]), "n_resident": n_resident}
            torch.save(ids, self.operations_dir / f"{i}.pt")
            if op == "query":
                q = self.queries[ids] if self.queries is not None else self.base_vectors[ids]
                t0 = time.time()
                resident_ids = self.all_ids[self.resident_set]
                gt_i, gt_d = knn(q, self.base_vectors[resident_ids], 100, self.metric)
                entry["gt_time"] = time.time() - t0
                torch.save(resident_ids[gt_i], self.operations_dir / f"{i}_gt_ids.pt")
                torch.save(gt_d, self.operations_dir / f"{i}_gt_dists.pt")
            self.runbook["operations"][i] = entry
            frac = np.zeros(all_sizes.shape[0])
            ru, rc = torch.unique(self.assignments[self.resident_set], return_counts=True)
            frac[ru] = (rc.float() / all_sizes[ru]).numpy()
            self.resident_history.append(frac)
        self.runbook["summary"] = {"n_inserts": counts["insert"], "n_deletes": counts["delete"],
                                   "n_queries": counts["query"], "n_operations": n_operations}
        np.save(self.workload_dir / "resident_history.npy", np.array(self.resident_history).T)
        with open(self.workload_dir / "runbook.json", "w") as f:
            json.dump(self.runbook, f, indent=4)
        return self.runbook


# ------------------------------------------------------------------------------------------------ evaluator
class WorkloadEvaluator:
    """workload_generator.py:401-606: build / load the initial index, replay the runbook, collect per-operation
    latency, recall and index state."""

    def __init__(self, workload_dir, output_dir, base_vectors_path=None):
        self.workload_dir, self.output_dir = to_path(workload_dir), to_path(output_dir)
        self.runbook_path = self.workload_dir / "runbook.json"
        self.operations_dir = self.workload_dir / "operations"
        self.initial_indices_path = self.workload_dir / "initial_indices.pt"
        self.base_vectors_path = to_path(base_vectors_path) if base_vectors_path else self.workload_dir / "base_vectors.pt"
        self.runbook = None

    def initialize_index(self, name, index, build_params, m_params):
        index_dir = self.workload_dir / "init_indexes"
        index_dir.mkdir(parents=True, exist_ok=True)
        index_path = index_dir / f"{name}.index"
        if not index_path.exists():
            vectors = torch.load(self.base_vectors_path, weights_only=True).to(torch.float32)
            initial = torch.load(self.initial_indices_path, weights_only=True).to(torch.int64)
            index.build(vectors[initial], ids=initial, **build_params)
            index.save(index_path)
        else:
            index.load(index_path, n_workers=build_params.get("num_workers", 0))
        if isinstance(index, QuakeWrapper) and m_params is not None:
            index.index.initialize_maintenance_policy(m_params)
        return index

    def evaluate_workload(self, name, index, build_params, search_params, do_maintenance=False, m_params=None,
                          batch=False) -> list:
        assert "k" in search_params, "search_params must contain 'k' for number of neighbors"
        base = torch.load(self.base_vectors_path, weights_only=True).to(torch.float32)
        index = self.initialize_index(name, index, build_params, m_params)
        self.runbook = json.load(open(self.runbook_path))
        queries = base if self.runbook["parameters"]["sample_queries"] else \
            torch.load(self.workload_dir / "query_vectors.pt", weights_only=True).to(torch.float32)
        self.runbook["initialize"]["time"] = 0.0
        sync = torch.cuda.synchronize if torch.cuda.is_available() else (lambda: None)
        results = []
        for op_id, op in self.runbook["operations"].items():
            kind = op["type"]
            ids = torch.load(self.operations_dir / f"{op_id}.pt", weights_only=True)
            recall = None
            sync()
            t0 = time.time()
            if kind == "insert":
                index.add(base[ids], ids=ids, num_threads=16)
            elif kind == "delete":
                index.remove(ids)
            elif kind == "query":
                q = queries[ids]
                if batch:
                    pred = index.search(q, **search_params).ids
                else:
                    pred = torch.cat([index.search(x.unsqueeze(0), **search_params).ids for x in q])
            sync()
            op_time = time.time() - t0
            if kind == "query":
                gt = torch.load(self.operations_dir / f"{op_id}_gt_ids.pt", weights_only=True)
                recall = float(compute_recall(pred, gt, search_params["k"]).mean())
                self.runbook["operations"][op_id]["recall"] = recall
            m_info = index.maintenance() if do_maintenance else None
            row = {"operation_number": int(op_id), "operation_type": kind, "latency_ms": op_time * 1000,
                   "recall": recall, "n_resident": op.get("n_resident")}
            if m_info is not None:
                row["maintenance_ms"] = getattr(m_info, "total_time_us", 0) / 1000.0
                row["n_splits"], row["n_deletes"] = getattr(m_info, "n_splits", 0), getattr(m_info, "n_deletes", 0)
            row.update(index.index_state())
            row.update(search_params)
            results.append(row)
        self.output_dir.mkdir(parents=True, exist_ok=True)
        with open(self.output_dir / f"{name}_results.json", "w") as f:
            json.dump(results, f, indent=1)
        self.summary = summarize(results)
        self._plot(results)
        return results

    def _plot(self, results) -> None:
        try:
            import matplotlib
            matplotlib.use("Agg")
            import matplotlib.pyplot as plt
        except Exception:
            return  # plots are optional
        fig, axs = plt.subplots(2, 2, figsize=(12, 10))
        for kind, mark in (("insert", "o"), ("delete", "s"), ("query", "^")):
            xs = [r["operation_number"] for r in results if r["operation_type"] == kind]
            if xs:
                axs[0, 0].plot(xs, [r["latency_ms"] for r in results if r["operation_type"] == kind], label=kind, marker=mark)
        axs[0, 0].set_xlabel("Operation Number"); axs[0, 0].set_ylabel("Latency (ms)"); axs[0, 0].legend()
        axs[0, 1].plot([r["operation_number"] for r in results], [r["n_list"] for r in results], marker="o")
        axs[0, 1].set_ylabel("Number of Partitions")
        axs[1, 0].plot([r["operation_number"] for r in results], [r["n_resident"] for r in results], marker="o")
        axs[1, 0].set_ylabel("Resident Vectors")
        qs = [r for r in results if r["recall"] is not None]
        axs[1, 1].plot([r["operation_number"] for r in qs], [r["recall"] for r in qs], marker="o")
        axs[1, 1].set_ylabel("Query Recall")
        plt.tight_layout()
        plt.savefig(self.output_dir / "evaluation_plots.png")
        plt.close()


def summarize(results) -> dict:
    """Mean latency per operation type and mean recall (the summary the reference prints)."""
    out = {}
    for kind in ("insert", "delete", "query"):
        lat = [r["latency_ms"] for r in results if r["operation_type"] == kind]
        out[f"avg_latency_{kind}_ms"] = float(np.mean(lat)) if lat else None
    rec = [r["recall"] for r in results if r["recall"] is not None]
    out["avg_query_recall"] = float(np.mean(rec)) if rec else None
    return out
