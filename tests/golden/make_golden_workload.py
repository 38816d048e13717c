"""Golden fixture for the workload harness: runs the REFERENCE's own DynamicWorkloadGenerator
(/root/reference/src/python/workload_generator.py, imported from where it lies, on top of the compiled unmodified
reference in oracle/_ref) on a small seeded dataset and records what it drew. Run in the build container:

    python tests/golden/make_golden_workload.py

tests/test_workload.py replays quake_b200.workload.DynamicWorkloadGenerator on the same inputs, with the reference's
clustering (assignments + centroids + the state it leaves torch's global generator in, recorded here) injected, and requires the same stream of operations: types, sampled
ids, resident-set sizes and ground-truth neighbours. The reference's Python package is assembled in memory (module
`quake` = the compiled bindings, sub-modules resolved from the reference tree); matplotlib, which it imports for plots,
is stubbed. Nothing of the reference is copied into this repository -- only its outputs."""
import json
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_PY = "/root/reference/src/python"
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
import quake_ref  # noqa: E402  the compiled, unmodified reference (oracle/build_ref.sh)

pkg = types.ModuleType("quake")
pkg.__path__ = [REF_PY]
for name in dir(quake_ref):
    if not name.startswith("_"):
        setattr(pkg, name, getattr(quake_ref, name))
sys.modules["quake"] = pkg
from unittest import mock  # noqa: E402
plt = mock.MagicMock()  # the generator draws a plot of the resident set at the end: every call is swallowed
plt.subplots.return_value = (mock.MagicMock(), mock.MagicMock())
mpl = types.ModuleType("matplotlib")
mpl.pyplot = plt
sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt
import quake.workload_generator as refgen  # noqa: E402


def dataset(seed, n, nq, d):
    g = torch.Generator().manual_seed(seed)
    centers = torch.randn(12, d, generator=g) * 4.0
    base = centers[torch.randint(0, 12, (n,), generator=g)] + torch.randn(n, d, generator=g)
    queries = centers[torch.randint(0, 12, (nq,), generator=g)] + torch.randn(nq, d, generator=g)
    return base, queries


CASES = {
    # name: (dataset seed, n, nq, d, generator kwargs)
    "skewed": (101, 3000, 200, 16, dict(metric="l2", insert_ratio=0.3, delete_ratio=0.2, query_ratio=0.5,
                                        update_batch_size=100, query_batch_size=20, number_of_operations=14,
                                        initial_size=1200, cluster_size=250, cluster_sample_distribution="skewed",
                                        query_cluster_sample_distribution="skewed", seed=77)),
    "uniform": (202, 2000, 150, 8, dict(metric="ip", insert_ratio=0.4, delete_ratio=0.3, query_ratio=0.3,
                                        update_batch_size=80, query_batch_size=16, number_of_operations=12,
                                        initial_size=900, cluster_size=200, cluster_sample_distribution="uniform",
                                        query_cluster_sample_distribution="uniform", seed=5)),
}

out = {}
meta = {}
for name, (ds, n, nq, d, kw) in CASES.items():
    base, queries = dataset(ds, n, nq, d)
    with tempfile.TemporaryDirectory() as tmp:
        gen = refgen.DynamicWorkloadGenerator(workload_dir=tmp, base_vectors=base, queries=queries, **kw)
        # the reference's clustering (its C++ k-means) draws from torch's global generator; what the generator itself
        # draws afterwards depends on that state, so it is part of the recorded clustering
        inner = gen.initialize_clustered_index

        def recording():
            index = inner()
            out[f"{name}_rng_after_clustering"] = torch.get_rng_state().numpy().copy()
            return index

        gen.initialize_clustered_index = recording
        gen.generate_workload()
        book = json.load(open(os.path.join(tmp, "runbook.json")))
        out[f"{name}_assignments"] = gen.assignments.numpy().astype(np.int64)
        out[f"{name}_centroids"] = gen.clustered_index.centroids().numpy().astype(np.float32)
        out[f"{name}_initial"] = torch.load(os.path.join(tmp, "initial_indices.pt")).numpy().astype(np.int64)
        ops = book["operations"]
        types_, sizes, resident, ids_all, offs = [], [], [], [], [0]
        for i in sorted(ops, key=int):
            e = ops[i]
            types_.append(e["type"])
            sizes.append(e["sample_size"])
            resident.append(e["n_resident"])
            ids = torch.load(os.path.join(tmp, "operations", f"{i}.pt")).numpy().astype(np.int64)
            ids_all.append(ids)
            offs.append(offs[-1] + len(ids))
            if e["type"] == "query":
                out[f"{name}_gt_{i}"] = torch.load(os.path.join(tmp, "operations", f"{i}_gt_ids.pt")).numpy().astype(np.int64)[:, :10]
        out[f"{name}_op_ids"] = np.concatenate(ids_all)
        out[f"{name}_op_offsets"] = np.array(offs, dtype=np.int64)
        out[f"{name}_n_resident"] = np.array(resident, dtype=np.int64)
        meta[name] = {"dataset": [ds, n, nq, d], "kwargs": kw, "types": types_, "sizes": sizes, "summary": book["summary"],
                      "parameters": book["parameters"]}
np.savez_compressed(os.path.join(HERE, "workload.npz"), **out)
json.dump(meta, open(os.path.join(HERE, "workload.json"), "w"), indent=1)
print({k: (v["types"], v["summary"]) for k, v in meta.items()})
print("wrote", os.path.join(HERE, "workload.npz"), os.path.getsize(os.path.join(HERE, "workload.npz")), "bytes")
