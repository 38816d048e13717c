"""Run the BASELINE.json configs other than the bench's headline one (C2) at full size on the GPU box and check the
size-independent properties that stand in for an oracle there; writes gpurun_out/configs_<tag>.json.

    python scripts/run_configs.py c1 c3 c5 [--scale 1.0] [--tag r1]
    torchrun ... scripts/run_configs.py c4 [--scale 1.0]          (lists sharded over the ranks)

C1 10k x 128, nlist 1024, nprobe 10, k 10, l2     -- the reference builds + saves, we load: ids/distances identical
C3 10M x 128 ip, nlist 16384, APS recall 0.9, k 100 -- recall vs brute force, partitions scanned, oracle on 8 queries
C4 100M x 96, nlist 65536, nprobe 64, k 10, l2, sharded -- every rank gets the same answer; self-queries found
C5 10M x 128 dynamic: +1M add, -100k remove, refit, search k 10 -- counts, removed ids gone, added vectors found
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def sync():
    torch.cuda.synchronize()


def timed(fn, reps=1):
    sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    sync()
    return out, (time.perf_counter() - t0) / reps


def brute_force(x_dev, q_dev, k, metric, chunk=1 << 20):
    """Exact top-k ids by chunked matmul on the device (checker only)."""
    best_d, best_i = None, None
    qn = (q_dev * q_dev).sum(1, keepdim=True)
    for s in range(0, x_dev.shape[0], chunk):
        xb = x_dev[s:s + chunk]
        ip = q_dev @ xb.T
        sc = -ip if metric == "ip" else (qn + (xb * xb).sum(1)[None, :] - 2 * ip)
        d, i = sc.topk(min(k, xb.shape[0]), largest=False)
        i = i + s
        if best_d is None:
            best_d, best_i = d, i
        else:
            d = torch.cat([best_d, d], 1)
            i = torch.cat([best_i, i], 1)
            sel = d.topk(k, largest=False).indices
            best_d, best_i = d.gather(1, sel), i.gather(1, sel)
    return best_i


def recall(ids, gt):
    hit = 0
    for a, b in zip(ids.tolist(), gt.tolist()):
        hit += len(set(a) & set(b))
    return hit / float(gt.numel())


def c1(args):
    import quake_b200 as qb
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    import quake_ref
    torch.manual_seed(1234)
    x = torch.randn(10000, 128)
    ids = torch.arange(10000, dtype=torch.int64)
    bp = quake_ref.IndexBuildParams(); bp.nlist, bp.metric, bp.niter = 1024, "l2", 5
    ref = quake_ref.QuakeIndex(); ref.build(x, ids, bp)
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        ref.save(os.path.join(tmp, "i"))
        idx = qb.QuakeIndex(); idx.load(os.path.join(tmp, "i"))
    for Q in (100, 1024):
        torch.manual_seed(4321)
        q = torch.randn(Q, 128)
        rsp = quake_ref.SearchParams(); rsp.k, rsp.nprobe = 10, 10
        sp = qb.SearchParams(); sp.k, sp.nprobe = 10, 10
        want = ref.search(q, rsp)
        for _ in range(8):  # captures the search plan for this batch size; lets the filter-precision policy settle (it may
            idx.search(q, sp)  # re-capture the plan once after its first 256 queries)
        got, t = timed(lambda: idx.search(q, sp), reps=20)
        _, tref = timed(lambda: ref.search(q, rsp), reps=3)
        out[f"Q{Q}"] = {"ids_equal_reference": bool(torch.equal(got.ids, want.ids)),
                        "dist_bit_equal_reference": bool(torch.equal(got.distances, want.distances)),
                        "e2e_us_per_batch": t * 1e6, "reference_us_per_batch": tref * 1e6}
    return out


def c3(args):
    import quake_b200 as qb
    from oracle import oracle as orc
    n = int(10_000_000 * args.scale); d = 128; nlist = max(64, int(16384 * args.scale)); Q = 1024; k = 100
    torch.manual_seed(1234)
    x = torch.randn(n, d); x /= x.norm(dim=1, keepdim=True)
    torch.manual_seed(4321)
    q = torch.randn(Q, d); q /= q.norm(dim=1, keepdim=True)
    bp = qb.IndexBuildParams(); bp.nlist, bp.metric, bp.niter = nlist, "ip", 5
    idx = qb.QuakeIndex()
    _, tb = timed(lambda: idx.build(x, torch.arange(n, dtype=torch.int64), bp))
    sp = qb.SearchParams(); sp.k, sp.recall_target, sp.initial_search_fraction = k, 0.9, 0.02
    idx.search(q[:64], sp)
    _, t_first = timed(lambda: idx.search(q, sp))  # first full batch: allocator growth (round buffers of several 100 MB)
    res, ts = timed(lambda: idx.search(q, sp), reps=3)
    scanned = idx.last_partitions_scanned.float()
    dev = idx.store.device
    gt = brute_force(x.to(dev), q.to(dev), k, "ip").cpu()
    out = {"n": n, "nlist": nlist, "build_s": tb, "search_s_1024q": ts, "first_search_s_1024q": t_first, "qps": Q / ts, "recall_at_100": recall(res.ids, gt),
           "mean_partitions_scanned": float(scanned.mean()), "max_partitions_scanned": float(scanned.max()),
           "candidates_per_query": max(int(nlist * 0.02), 1)}
    # fixed nprobe at the same k for comparison (the non-adaptive hot path)
    sp2 = qb.SearchParams(); sp2.k, sp2.nprobe = k, 64
    idx.search(q, sp2)
    r2, t2 = timed(lambda: idx.search(q, sp2), reps=3)
    out["fixed_nprobe64"] = {"search_s_1024q": t2, "qps": Q / t2, "recall_at_100": recall(r2.ids, gt)}
    # oracle on 8 queries, same index content
    oi, od, _cid, sc = orc.search_lists(*_lists(idx, orc), q[:8], k, 0, idx.metric, recall_target=0.9,
                                        initial_search_fraction=0.02, return_probe=True)
    r8 = idx.search(q[:8], sp)
    out["oracle_8q"] = {"ids_equal": bool(torch.equal(r8.ids, oi)), "dist_bit_equal": bool(torch.equal(r8.distances, od)),
                        "scanned_equal": bool(np.array_equal(idx.last_partitions_scanned.cpu().numpy(), sc))}
    return out


def _lists(idx, orc):
    pids, lists, cv, ci = orc.index_lists(idx)
    return pids, lists, cv, ci


def c5(args):
    import quake_b200 as qb
    n0 = int(10_000_000 * args.scale); n_add = int(1_000_000 * args.scale); n_rm = int(100_000 * args.scale)
    d = 128; nlist = max(64, int(16384 * args.scale)); Q = 1024
    torch.manual_seed(1234)
    x = torch.randn(n0, d)
    bp = qb.IndexBuildParams(); bp.nlist, bp.metric, bp.niter = nlist, "l2", 5
    idx = qb.QuakeIndex()
    _, tb = timed(lambda: idx.build(x, torch.arange(n0, dtype=torch.int64), bp))
    torch.manual_seed(77)
    xa = torch.randn(n_add, d)
    ida = torch.arange(n0, n0 + n_add, dtype=torch.int64)
    _, ta = timed(lambda: idx.add(xa, ida))
    g = torch.Generator().manual_seed(99)
    rm = torch.randperm(n0 + n_add, generator=g)[:n_rm].to(torch.int64)
    _, tr = timed(lambda: idx.remove(rm))
    out = {"n0": n0, "build_s": tb, "add_s": ta, "add_vectors_per_s": n_add / ta, "remove_s": tr,
           "ntotal_ok": idx.ntotal() == n0 + n_add - n_rm}
    # maintenance refit: the partitions touched by a window of queries (refine_partitions, 3 iterations)
    torch.manual_seed(4321)
    q = torch.randn(Q, d)
    sp = qb.SearchParams(); sp.k, sp.nprobe = 10, 64
    idx.search(q, sp)  # fills the hit window (window_size 1000)
    _, tm = timed(lambda: idx.maintenance())
    out["maintenance_s"] = tm
    out["ntotal_after_refit_ok"] = idx.ntotal() == n0 + n_add - n_rm
    idx.search(q, sp)
    res, ts = timed(lambda: idx.search(q, sp), reps=5)
    out["search_s_1024q"] = ts; out["qps"] = Q / ts
    rm_set = set(rm.tolist())
    out["removed_ids_never_returned"] = not any(int(v) in rm_set for v in res.ids.reshape(-1).tolist())
    # added vectors that were not removed again are their own nearest neighbour at distance 0
    keep = [i for i in range(min(n_add, 4096)) if int(ida[i]) not in rm_set][:512]
    sp1 = qb.SearchParams(); sp1.k, sp1.nprobe = 1, 8
    r1 = idx.search(xa[keep], sp1)
    out["added_vectors_found"] = float((r1.ids[:, 0] == ida[keep]).float().mean())
    out["added_vectors_zero_distance"] = float((r1.distances[:, 0] == 0).float().mean())
    return out


def c4(args):
    import torch.distributed as dist
    import quake_b200 as qb
    from quake_b200.sharded import ShardedQuakeIndex
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    n = int(100_000_000 * args.scale); d = 96; nlist = max(64, int(65536 * args.scale)); Q = 1024
    lo, hi = rank * n // world, (rank + 1) * n // world
    g = torch.Generator().manual_seed(1234 + rank)
    x = torch.randn(hi - lo, d, generator=g)
    ids = torch.arange(lo, hi, dtype=torch.int64)
    bp = qb.IndexBuildParams(); bp.nlist, bp.metric, bp.niter = nlist, "l2", 5
    sh = ShardedQuakeIndex()
    _, tb = timed(lambda: sh.build(x, ids, bp))
    torch.manual_seed(4321)
    q = torch.randn(Q, d)
    q[:256] = x[:256] if rank == 0 else q[:256]
    if world > 1:
        qd = q.cuda(); dist.broadcast(qd, 0); q = qd.cpu()
    sp = qb.SearchParams(); sp.k, sp.nprobe = 10, 64
    sh.search(q, sp)
    res, ts = timed(lambda: sh.search(q, sp), reps=5)
    same = True
    if world > 1:
        a = res.ids.cuda(); parts = [torch.zeros_like(a) for _ in range(world)]
        dist.all_gather(parts, a)
        same = all(torch.equal(p, parts[0]) for p in parts)
    out = {"n": n, "nlist": nlist, "world": world, "build_s": tb, "search_s_1024q": ts, "qps": Q / ts,
           "ntotal_ok": sh.ntotal() == n, "all_ranks_same_answer": bool(same),
           "self_queries_found": float((res.ids[:256, 0] == torch.arange(256)).float().mean()),
           "local_vectors_rank0": sh.local.store.ntotal}
    if world > 1:
        dist.barrier(); dist.destroy_process_group()
    return out if rank == 0 else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="+")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--tag", default="r1")
    args = ap.parse_args()
    results = {}
    for c in args.configs:
        t0 = time.perf_counter()
        try:
            r = {"c1": c1, "c3": c3, "c4": c4, "c5": c5}[c](args)
        except Exception as e:  # keep going: one config must not hide the others
            import traceback
            r = {"error": repr(e), "trace": traceback.format_exc()[-1500:]}
        if r is not None:
            r["wall_s"] = time.perf_counter() - t0
            r["scale"] = args.scale
            results[c] = r
            print(c, json.dumps(r), flush=True)
    if results:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        name = "configs_%s_%s.json" % (args.tag, "_".join(sorted(results)))
        json.dump(results, open(os.path.join(ROOT, "gpurun_out", name), "w"), indent=1)


if __name__ == "__main__":
    main()
