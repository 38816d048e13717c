"""Small-batch latency probe: C1-shaped (10k x 128, nlist 1024, nprobe 10) and C2-shaped indexes at Q in {1, 10, 100,
1024}; per batch size: e2e (host tensors, graph plan), device-resident eager, device-resident graph. Optional env
QK_PROBE_NCU=1 wraps ONE eager Q=100 search in cudaProfilerStart/Stop for an ncu launch list.

    python scripts/latency_probe.py [c1] [c2]
"""
from __future__ import annotations

import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import quake_b200 as qb  # noqa: E402
from quake_b200 import index as qi  # noqa: E402


def build(n, d, nlist):
    torch.manual_seed(1234)
    x = torch.randn(n, d)
    bp = qb.IndexBuildParams()
    bp.nlist, bp.metric, bp.niter = nlist, "l2", 5
    idx = qb.QuakeIndex()
    idx.build(x, torch.arange(n, dtype=torch.int64), bp)
    return idx


def timeit(fn, reps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e6


def probe(name, idx, nprobe, k=10):
    out = {}
    dev = idx.store.device
    for Q in (1, 10, 100, 1024):
        torch.manual_seed(4321)
        q = torch.randn(Q, idx.store.d)
        qp = q.pin_memory()
        qd = qi.clustering.pad_rows(q, dev)
        sp = qb.SearchParams()
        sp.k, sp.nprobe = k, nprobe
        r = {}
        qi.GRAPHS_ENABLED = False
        for _ in range(3):
            idx._search_device(qd, sp)
        r["device_eager_us"] = timeit(lambda: idx._search_device(qd, sp), 20)
        if os.environ.get("QK_PROBE_NCU") == "1" and Q == 100:
            torch.cuda.cudart().cudaProfilerStart()
            idx._search_device(qd, sp)
            torch.cuda.synchronize()
            torch.cuda.cudart().cudaProfilerStop()
        r["e2e_eager_us"] = timeit(lambda: idx.search(qp, sp), 20)
        qi.GRAPHS_ENABLED = True
        for _ in range(3):
            idx._search_device(qd, sp)
        r["device_graph_us"] = timeit(lambda: idx._search_device(qd, sp), 50)
        for _ in range(3):
            idx.search(qp, sp)
        r["e2e_graph_us"] = timeit(lambda: idx.search(qp, sp), 50)
        # pageable host input (what a reference user passes), call by call
        per = []
        for _ in range(20):
            t0 = time.perf_counter()
            idx.search(q, sp)
            per.append((time.perf_counter() - t0) * 1e6)
        r["e2e_pageable_median_us"] = sorted(per)[10]
        r["e2e_pageable_max_us"] = max(per)
        r["e2e_pageable_first_us"] = per[0]
        out[f"Q{Q}"] = {a: round(b, 1) for a, b in r.items()}
        print(name, Q, out[f"Q{Q}"], flush=True)
    return out


def build_by_reference(n, d, nlist, nq_ref):
    """C1 as scripts/run_configs.py runs it: the REFERENCE builds and saves, we load; the reference also searches in
    this process (its OpenMP / std::async threads stay around afterwards)."""
    import tempfile
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    import quake_ref
    torch.manual_seed(1234)
    x = torch.randn(n, d)
    bp = quake_ref.IndexBuildParams()
    bp.nlist, bp.metric, bp.niter = nlist, "l2", 5
    ref = quake_ref.QuakeIndex()
    ref.build(x, torch.arange(n, dtype=torch.int64), bp)
    with tempfile.TemporaryDirectory() as tmp:
        ref.save(os.path.join(tmp, "i"))
        idx = qb.QuakeIndex()
        idx.load(os.path.join(tmp, "i"))
    return idx, ref, quake_ref


def main():
    which = sys.argv[1:] or ["c1", "c2"]
    res = {}
    if "c1ref" in which:
        idx, ref, quake_ref = build_by_reference(10000, 128, 1024, 100)
        sizes = idx.store.list_size
        print("c1ref lists:", len(sizes), "empty:", int((sizes == 0).sum()), "max:", int(sizes.max()), flush=True)
        res["c1ref_before_ref_search"] = probe("c1ref/before", idx, 10)
        for Q in (100, 1024):
            torch.manual_seed(4321)
            q = torch.randn(Q, 128)
            rsp = quake_ref.SearchParams()
            rsp.k, rsp.nprobe = 10, 10
            t0 = time.perf_counter()
            for _ in range(3):
                ref.search(q, rsp)
            print("reference search us", Q, (time.perf_counter() - t0) / 3 * 1e6, flush=True)
        res["c1ref_after_ref_search"] = probe("c1ref/after", idx, 10)
        print("torch threads", torch.get_num_threads(), "OMP", os.environ.get("OMP_NUM_THREADS"), flush=True)
    if "c1" in which:
        res["c1"] = probe("c1", build(10000, 128, 1024), 10)
    if "c2" in which:
        res["c2"] = probe("c2", build(1_000_000, 128, 4096), 64)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    tag = os.environ.get("QK_PROBE_TAG", "r2")
    with open(os.path.join(ROOT, "gpurun_out", f"latency_{tag}.json"), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
