#!/usr/bin/env python
"""bench.py -- the Quake search hot path on B200 (BASELINE.json metric: QPS, k=10, d=128).

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA kernels behind the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU search, same workload

Workload at every N: BASELINE.json configs[1] -- 1M x 128 float32 (torch.randn, seed 1234), nlist=4096,
Q=1024 queries (seed 4321), k=10, l2, nprobe=64 (the headline nprobe of SURVEY.md 8d). A "step" is one
search of the 1024-query batch: coarse centroid scan -> partition scan -> top-k. Multi-GPU (N>1): one
replica of the index per rank and an independent 1024-query batch per rank ("replicas only" for this
config; weak scaling), no data-path collective; the line additionally carries a `sharded` object --
the one split the north star names (lists sharded over the ranks + all-gather of the partial top-k), on a
C4-shaped workload sized to the box.

value   = queries/s with the queries already resident in HBM (CUDA events around K steps, max over ranks; the K-step
          region is repeated until >= 100 ms have been timed and the MEDIAN repeat is reported)
e2e     = queries/s through QuakeIndex.search with pinned HOST query tensors in and host results out (same protocol)
roofline= the partition-scan filter kernel: algorithmic bytes (every distinct probed list read once,
          n_p * d * 4 B) / its CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs
cpu_baseline = the compiled reference (oracle/_ref) searching the same index on the host cores
parity  = ALL queries of the batch against the compiled reference on the same index (ids, distances)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = {"name": "C2: 1M x 128 f32 randn, nlist=4096, Q=1024, k=10, l2, nprobe=64",
            "N": 1_000_000, "d": 128, "nlist": 4096, "Q": 1024, "k": 10, "nprobe": 64, "metric": "l2", "niter": 5}
METRIC = "search_qps_k10_d128"
UNIT = "queries/s"
MIN_TIMED_MS = 100.0  # every reported phase times at least this much GPU work (repeats of the K-step region)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nprobe", type=int, default=WORKLOAD["nprobe"])
    ap.add_argument("--n", type=int, default=WORKLOAD["N"], help="override N (debug only; the line says so)")
    ap.add_argument("--nlist", type=int, default=WORKLOAD["nlist"])
    ap.add_argument("--q", type=int, default=WORKLOAD["Q"], help="override the batch size (debug only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="headline numbers only (no latency / sweep / build sections)")
    return ap.parse_args()


def workload_name(W):
    tag = "C2" if (W["N"], W["nlist"], W["Q"], W["nprobe"]) == (1_000_000, 4096, 1024, 64) else "modified (debug)"
    return (f"{tag}: {W['N']} x {W['d']} f32 randn, nlist={W['nlist']}, Q={W['Q']}, k={W['k']}, {W['metric']}, "
            f"nprobe={W['nprobe']}")


def config_of(W, gpus):
    """The `config` object -- identical in both arms for the same command line."""
    return {"workload": W["name"], "N": W["N"], "nlist": W["nlist"], "nprobe": W["nprobe"], "Q": W["Q"], "k": W["k"],
            "metric": W["metric"], "parallelism": f"replicas x{gpus}" if gpus > 1 else "1 gpu",
            "l2_policy": "index (512 MB of lists) larger than L2 (126 MB); no flush between steps"}


def make_data(n, d, q, rank=0):
    import torch
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(n, d, generator=g)
    g2 = torch.Generator().manual_seed(4321 + rank)
    xq = torch.randn(q, d, generator=g2)
    return x, xq


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [t.strip() for t in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ reference (CPU)
def import_reference():
    p = os.path.join(ROOT, "oracle", "_ref")
    if p not in sys.path:
        sys.path.insert(0, p)
    import quake_ref  # the compiled, unmodified reference (built by oracle/build_ref.sh)
    return quake_ref


def reference_build_note():
    return ("compiled from /root/reference by oracle/build_ref.sh with -O3 -march=x86-64-v3 (AVX2+FMA; the "
            "reference's own CMake uses -march=native), num_workers=0 (no worker/NUMA path)")


def time_reference_search(ref_idx, quake_ref, xq, k, nprobe, cores, budget_s, steps=None, warmup=1):
    """Times the reference's CPU search (num_workers=0) in its two deterministic modes and returns the
    faster: serial_scan parallel over queries (num_threads=cores) and batched_serial_scan."""
    import torch
    torch.set_num_threads(cores)
    best = None
    for mode, batched in (("serial_scan", False), ("batched_serial_scan", True)):
        sp = quake_ref.SearchParams()
        sp.k, sp.nprobe, sp.batched_scan, sp.num_threads = k, nprobe, batched, cores
        t0 = time.perf_counter()
        ref_idx.search(xq, sp)  # warm-up + cost probe
        one = time.perf_counter() - t0
        reps = steps if steps is not None else max(1, min(10, int(budget_s / 2 / max(one, 1e-3))))
        for _ in range(max(0, warmup - 1)):
            ref_idx.search(xq, sp)
        t0 = time.perf_counter()
        for _ in range(reps):
            ref_idx.search(xq, sp)
        dt = (time.perf_counter() - t0) / reps
        qps = xq.shape[0] / dt
        if best is None or qps > best["value"]:
            best = {"value": qps, "mode": mode, "ms_per_step": dt * 1e3, "reps": reps}
    return best


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path, all host threads, same config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    W = dict(WORKLOAD, N=args.n, nlist=args.nlist, nprobe=args.nprobe, Q=args.q)
    W["name"] = workload_name(W)
    cores = host_cores()
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    line = {"impl": "reference", "metric": METRIC, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic"}
    try:
        quake_ref = import_reference()
        kind = "reference"
    except Exception as e:  # the compiled reference did not travel / load: fall back to the oracle port
        quake_ref, kind, why = None, "port", repr(e)
    x, xq = make_data(W["N"], W["d"], W["Q"])
    ids = torch.arange(W["N"], dtype=torch.int64)
    if kind == "reference":
        torch.set_num_threads(cores)
        bp = quake_ref.IndexBuildParams()
        bp.nlist, bp.metric, bp.niter = W["nlist"], W["metric"], W["niter"]
        idx = quake_ref.QuakeIndex()
        t0 = time.perf_counter()
        idx.build(x, ids, bp)
        build_s = time.perf_counter() - t0
        # bound the run: probe one step, shrink the per-step sample if K+W steps would not end in ~3 min
        sp = quake_ref.SearchParams()
        sp.k, sp.nprobe, sp.num_threads = W["k"], W["nprobe"], cores
        t0 = time.perf_counter(); idx.search(xq, sp); one = time.perf_counter() - t0
        total_steps = args.steps + args.warmup
        qn = W["Q"]
        while qn > 32 and one * (qn / W["Q"]) * total_steps * 2 > 180:
            qn //= 2
        best = time_reference_search(idx, quake_ref, xq[:qn], W["k"], W["nprobe"], cores, 0, steps=args.steps,
                                     warmup=max(args.warmup, 1))
        sample = (f"{qn} of {W['Q']} queries per step, {best['mode']}, num_threads={cores}, index built by the reference "
                  f"in {build_s:.1f}s; {reference_build_note()}")
    else:
        from oracle import oracle as orc
        import numpy as np
        # scalar port on a small index slice (bounded): 64 queries, lists from a 100k subset
        n_small = min(W["N"], 100_000)
        xs = x[:n_small].numpy()
        cents = xs[: max(W["nlist"] // 10, 1)]
        a = orc.assign(xs, cents, W["metric"])
        lists = [(xs[a == c], np.nonzero(a == c)[0].astype(np.int64)) for c in range(cents.shape[0])]
        qn = 64
        t0 = time.perf_counter()
        for _ in range(args.steps):
            orc.search_lists(np.arange(len(lists)), lists, cents, np.arange(len(lists)), xq[:qn].numpy(), W["k"],
                             min(W["nprobe"], len(lists)), W["metric"])
        dt = (time.perf_counter() - t0) / args.steps
        best = {"value": qn / dt, "ms_per_step": dt * 1e3, "mode": "oracle port"}
        cores = 1
        sample = f"oracle port (compiled reference unavailable: {why}); {qn} queries, {n_small} vectors"
    line.update({"value": best["value"], "ms_per_step": best["ms_per_step"], "config": config_of(W, args.gpus),
                 "cpu_baseline": {"value": best["value"], "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
                 "e2e": {"value": best["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                 "gpu_launches": 0})
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------ this repo (GPU)
def log(msg):
    """Progress on stderr (rank-tagged): a hung multi-rank run shows where it stopped."""
    print(f"[bench rank {os.environ.get('RANK', '0')} +{time.perf_counter() - _T0:.1f}s] {msg}", file=sys.stderr, flush=True)


_T0 = time.perf_counter()


def median(v):
    v = sorted(v)
    return v[len(v) // 2]


def compare_with_reference(got_ids, got_dist, want_ids, want_dist):
    """BASELINE.md section 5 / SURVEY 8d parity gate over ALL queries: ids equal, distances bit-equal, id mismatches
    that are swaps of reference near-ties (gap < 1e-6 relative) counted separately."""
    import numpy as np
    gi, gd, wi, wd = got_ids.numpy(), got_dist.numpy(), want_ids.numpy(), want_dist.numpy()
    bad = np.argwhere(gi != wi)
    near = 0
    for q, j in bad:
        nb = [t for t in (j - 1, j + 1) if 0 <= t < wi.shape[1]]
        if any(abs(wd[q, j] - wd[q, t]) <= 1e-6 * max(abs(wd[q, j]), 1e-30) for t in nb):
            near += 1
    fin = np.isfinite(wd)
    rel = np.abs(gd[fin] - wd[fin]) / np.maximum(np.abs(wd[fin]), 1e-30)
    return {"queries": int(gi.shape[0]), "ids_equal": bool(len(bad) == 0), "id_mismatches": int(len(bad)),
            "id_mismatches_that_are_near_tie_swaps": int(near),
            "distances_bit_equal": int((gd[fin] == wd[fin]).sum()), "distances_compared": int(fin.sum()),
            "max_rel_distance_error": float(rel.max()) if rel.size else 0.0, "tolerance": 1e-4}


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; quake_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    import quake_b200 as qb
    from quake_b200 import _lib
    from quake_b200 import index as _qi
    import ctypes as C
    lib = _lib.load()

    W = dict(WORKLOAD, N=args.n, nlist=args.nlist, nprobe=args.nprobe, Q=args.q)
    W["name"] = workload_name(W)
    if os.environ.get("QK_BENCH_ONLY_SHARDED") == "1" and world > 1:  # debugging aid: just the sharded section
        out = sharded_section(qb, world, rank, dev, args)
        if rank == 0:
            print(json.dumps({"sharded": out}), flush=True)
        dist.barrier()
        dist.destroy_process_group()
        return 0
    x, xq_h = make_data(W["N"], W["d"], W["Q"], rank)
    ids = torch.arange(W["N"], dtype=torch.int64)
    bp = qb.IndexBuildParams()
    bp.nlist, bp.metric, bp.niter = W["nlist"], W["metric"], W["niter"]
    idx = qb.QuakeIndex()
    t0 = time.perf_counter()
    binfo = idx.build(x, ids, bp)
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0

    sp = qb.SearchParams()
    sp.k, sp.nprobe = W["k"], W["nprobe"]
    xq_pinned = xq_h.pin_memory()
    xq_d = xq_h.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- the compiled reference on the SAME index (our save -> its load): parity over all queries now, cpu_baseline later
    ref = quake_ref = None
    ref_why = None
    if rank == 0 and not args.no_cpu_baseline:
        try:
            quake_ref = import_reference()
            with tempfile.TemporaryDirectory() as tmp:
                p = os.path.join(tmp, "idx")
                idx.save(p)
                ref = quake_ref.QuakeIndex()
                ref.load(p, 0)
        except Exception as e:
            ref, ref_why = None, repr(e)

    log(f"index built in {build_s:.2f}s")
    parity = {}
    if rank == 0:
        full = idx.search(xq_h, sp)
        if ref is not None:
            rsp = quake_ref.SearchParams()
            rsp.k, rsp.nprobe, rsp.num_threads = W["k"], W["nprobe"], host_cores()
            want = ref.search(xq_h, rsp)
            parity["vs_compiled_reference"] = compare_with_reference(full.ids, full.distances, want.ids, want.distances)
            parity[f"ids_equal_reference_{W['Q']}q"] = parity["vs_compiled_reference"]["ids_equal"]
            parity["reference_build"] = reference_build_note()
        else:
            from oracle import oracle as orc
            oi, od = orc.search_index_like(idx, xq_h[:16], k=W["k"], nprobe=W["nprobe"])
            parity["ids_equal_oracle_port_16q"] = bool(torch.equal(full.ids[:16], oi))
            parity["dist_bit_equal_oracle_port_16q"] = bool(torch.equal(full.distances[:16], od))
            parity["compiled_reference_unavailable"] = ref_why or "--no-cpu-baseline"
        xd = x.to(dev)
        gt = torch.cdist(xq_d, xd).topk(W["k"], largest=False).indices.cpu()
        parity["recall_at_k_vs_bruteforce"] = recall_at_k(full.ids, gt)

    # ---- algorithmic bytes of the partition-scan filter kernel for this batch (grouped mode: every
    #      distinct probed list read once)
    psp = qb.SearchParams(); psp.k = min(W["nprobe"], idx.nlist()); psp.batched_scan = True
    p_ids, _, _ = idx.parent._search_device(xq_d, psp)
    sizes_all = torch.from_numpy(idx.store.list_size[[idx.store.pid_slot[int(p)] for p in idx.store.partition_ids()]])
    uniq = torch.unique(p_ids[p_ids >= 0]).cpu()
    alg_bytes = int(sizes_all[uniq].sum()) * W["d"] * 4
    pair_bytes = int(sizes_all[p_ids.cpu().reshape(-1)].sum()) * W["d"] * 4

    # ---- eager phases (no CUDA graph): selection statistics of one partition scan, then the per-launch timing of the
    #      filter kernel with the library's own CUDA-event pair around it (qk_profile_*)
    log("parity done; eager kernel timing")
    _qi.GRAPHS_ENABLED = False
    os.environ["QK_SCAN_STATS"] = "1"
    idx._search_device(xq_d, sp)
    torch.cuda.synchronize()
    st4 = _qi.LAST_SCAN_STATS.cpu().tolist()
    os.environ.pop("QK_SCAN_STATS")
    scan_stats = {"queries_rescanned": st4[0], "max_appended_per_query": st4[1],
                  "mean_appended_per_query": ((st4[3] << 32) | (st4[2] & 0xffffffff)) / W["Q"],
                  "threshold_refreshes": st4[6], "refresh_requests_dropped": st4[7]}
    for _ in range(args.warmup):
        idx._search_device(xq_d, sp)
    lib.qk_profile_begin(4 * args.steps + 8)
    barrier()
    cuprof = os.environ.get("QK_BENCH_CUPROF") == "1"  # ncu --profile-from-start off: capture these eager steps only
    if cuprof:
        torch.cuda.cudart().cudaProfilerStart()
    l0 = _qi.launch_count()
    for _ in range(args.steps):
        idx._search_device(xq_d, sp)
    barrier()
    launches_per_step = (_qi.launch_count() - l0) / args.steps
    if cuprof:
        torch.cuda.cudart().cudaProfilerStop()
    # filter-kernel records: 2 scan calls per step (coarse scan of the centroid list, then the partition scan)
    scan_ms = []
    for i in range(lib.qk_profile_count()):
        ms, qn, npb, kk = C.c_float(), C.c_int64(), C.c_int32(), C.c_int32()
        lib.qk_profile_read(i, C.byref(ms), C.byref(qn), C.byref(npb), C.byref(kk))
        if npb.value == W["nprobe"] and kk.value == W["k"]:
            scan_ms.append(ms.value)
    lib.qk_profile_end()
    scan_ms_avg = sum(scan_ms) / max(len(scan_ms), 1)
    _qi.GRAPHS_ENABLED = os.environ.get("QK_GRAPH", "1") != "0"

    # ---- device-resident timing (value): the step as the library runs it (CUDA-graph replay). The K-step region is
    #      repeated R times back to back (>= MIN_TIMED_MS of GPU work in total) and the median repeat is reported: one
    #      6 ms region is at the mercy of a single host hiccup (e.g. the clock sampler's nvidia-smi poll).
    log("device-resident timing")
    for _ in range(max(args.warmup, 3)):
        idx._search_device(xq_d, sp)
    sampler = ClockSampler(local_rank)
    sampler.start()
    t_wait = time.perf_counter()
    while len(sampler.rows) < 1 and time.perf_counter() - t_wait < 3.0:  # NVML start-up stays outside the timed regions
        for _ in range(10):
            idx._search_device(xq_d, sp)
        torch.cuda.synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        idx._search_device(xq_d, sp)
    e1.record()
    barrier()
    probe_ms = e0.elapsed_time(e1)
    repeats = agreed_repeats(probe_ms, world, dev)
    l0 = _qi.launch_count()
    dev_runs = []
    for _ in range(repeats):
        barrier()
        e0.record()
        for _ in range(args.steps):
            idx._search_device(xq_d, sp)
        e1.record()
        barrier()
        dev_runs.append(e0.elapsed_time(e1) / args.steps)
    timed_launches = _qi.launch_count() - l0
    dev_ms = median(dev_runs)

    # ---- end to end through the public API with host tensors (e2e), same protocol
    for _ in range(max(args.warmup, 3)):
        idx.search(xq_pinned, sp)
    e2e_runs = []
    l0 = _qi.launch_count()
    for _ in range(repeats):
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            r = idx.search(xq_pinned, sp)
        barrier()
        e2e_runs.append((time.perf_counter() - t0) * 1e3 / args.steps)
    timed_launches += _qi.launch_count() - l0
    clocks = sampler.stop()
    e2e_ms = median(e2e_runs)
    h2d = xq_pinned.numel() * 4
    d2h = r.ids.numel() * 8 + r.distances.numel() * 4

    # ---- max over ranks
    t = torch.tensor([dev_ms, e2e_ms, scan_ms_avg], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, scan_ms_avg = [float(v) for v in t.cpu()]

    extra = {}
    if not args.quick:
        if rank == 0:
            log("latency section")
            extra["latency"] = latency_section(qb, quake_ref, idx, ref, W, dev)
            log("nprobe sweep")
            extra["nprobe_sweep"] = nprobe_sweep(qb, idx, xq_d, xq_h, x, W, dev)
            log("build section")
            extra["build"] = build_section(qb, x, W, dev, binfo, build_s)
        if world > 1:
            del x
            log("waiting for the sharded section")
            barrier()
            try:
                extra["sharded"] = sharded_section(qb, world, rank, dev, args)
            except Exception as e:  # an error every rank hits alike (e.g. out of memory): keep the headline line
                extra["sharded"] = {"error": f"{type(e).__name__}: {e}"[:300]}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        achieved = alg_bytes / (scan_ms_avg * 1e-3) / 1e9 if scan_ms_avg > 0 else 0.0
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "scan_kernel_traffic.json"))).get("dram_bytes_per_launch")
        except Exception:
            pass
        terms = int(idx.store.filter_terms)
        line = {
            "metric": METRIC, "value": world * W["Q"] / (dev_ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(W, world),
            "timing": {"region": f"{args.steps} steps, repeated {repeats}x, median repeat reported",
                       "ms_per_step_min": min(dev_runs), "ms_per_step_max": max(dev_runs),
                       "e2e_ms_per_step_min": min(e2e_runs), "e2e_ms_per_step_max": max(e2e_runs)},
            "parity": parity, "scan_stats": scan_stats, "build_s": round(build_s, 2),
            "roofline": {"bound": "hbm", "kernel": "scan_mma_kernel (partition-scan filter, tcgen05)", "achieved": achieved,
                         "peak": peak, "unit": "GB/s", "frac": achieved / peak if peak else None, "traffic": traffic,
                         "traffic_source": "profiles/scan_kernel_traffic.json (one ncu --set full capture of this command)",
                         "traffic_over_algorithmic": round(traffic / alg_bytes, 4) if traffic and alg_bytes else None,
                         "filter": f"{terms}xTF32 tensor-core filter (3 = a_hi.b_hi + a_lo.b_hi + a_hi.b_lo, 2 = no a_lo term; chosen "
                                   "on evidence by the index, results exact either way: refine + proof + exact re-scan)",
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                         "per_query_list_bytes_per_launch": pair_bytes, "kernel_ms": scan_ms_avg,
                         "kernel_share_of_step": scan_ms_avg / dev_ms if dev_ms else None,
                         "step_hbm_frac": (alg_bytes / (dev_ms * 1e-3) / 1e9) / peak if peak and dev_ms else None},
            "e2e": {"value": world * W["Q"] / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(timed_launches), "gpu_launches_per_step": launches_per_step,
            "clocks": clocks,
        }
        line.update(extra)
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(idx, ref, quake_ref, xq_h, W, ref_why)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def agreed_repeats(probe_ms: float, world: int, device) -> int:
    """How many times the K-step region is repeated so that >= MIN_TIMED_MS of GPU work is timed -- THE SAME NUMBER ON
    EVERY RANK: each region is bracketed by collectives, and a rank that derived one repeat fewer from its own probe
    time would leave the others waiting in a barrier for ever (the 8-GPU run of round 2 hung exactly there). The probe
    time is all-reduced (max) first; the result is a function of that one number only."""
    if world > 1:
        import torch
        import torch.distributed as dist
        pm = torch.tensor([float(probe_ms)], dtype=torch.float64, device=device)
        dist.all_reduce(pm, op=dist.ReduceOp.MAX)
        probe_ms = float(pm.item())
    return int(min(60, max(3, -(-MIN_TIMED_MS // max(probe_ms, 1e-3)))))


def recall_at_k(ids, gt):
    hit = sum(len(set(a.tolist()) & set(b.tolist())) for a, b in zip(ids, gt))
    return hit / float(gt.numel())


def timed_us(fn, reps, sync):
    sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    sync()
    return (time.perf_counter() - t0) / reps * 1e6


def latency_section(qb, quake_ref, idx, ref, W, dev):
    """Small-batch latency, end to end (pageable host tensors in, host results out): Q in {1, 100} on C1 (10k x 128,
    nlist 1024, nprobe 10 -- examples/quickstart.py:31-68) and on this bench's index, the compiled reference beside
    it (same index, same queries, num_threads = 1 as in the quickstart)."""
    import torch
    out = {"unit": "us per batch, e2e, median of 50"}
    sync = torch.cuda.synchronize
    cases = []
    # C1: the reference builds (when available), we load
    torch.manual_seed(1234)
    x1 = torch.randn(10000, 128)
    ref1 = None
    idx1 = qb.QuakeIndex()
    if quake_ref is not None:
        bp = quake_ref.IndexBuildParams()
        bp.nlist, bp.metric, bp.niter = 1024, "l2", 5
        ref1 = quake_ref.QuakeIndex()
        ref1.build(x1, torch.arange(10000, dtype=torch.int64), bp)
        with tempfile.TemporaryDirectory() as tmp:
            ref1.save(os.path.join(tmp, "c1"))
            idx1.load(os.path.join(tmp, "c1"))
    else:
        bp = qb.IndexBuildParams()
        bp.nlist, bp.metric, bp.niter = 1024, "l2", 5
        idx1.build(x1, torch.arange(10000, dtype=torch.int64), bp)
    cases.append(("C1", idx1, ref1, 10))
    cases.append(("C2", idx, ref, W["nprobe"]))
    for name, ours, theirs, nprobe in cases:
        for Q in (1, 100):
            torch.manual_seed(4321)
            q = torch.randn(Q, 128)
            sp = qb.SearchParams()
            sp.k, sp.nprobe = 10, nprobe
            for _ in range(3):
                ours.search(q, sp)
            per = []
            for _ in range(50):
                t0 = time.perf_counter()
                ours.search(q, sp)
                per.append((time.perf_counter() - t0) * 1e6)
            row = {"gpu_us": round(median(per), 1)}
            if theirs is not None:
                rsp = quake_ref.SearchParams()
                rsp.k, rsp.nprobe = 10, nprobe
                want = theirs.search(q, rsp)
                got = ours.search(q, sp)
                row["ids_equal_reference"] = bool(torch.equal(got.ids, want.ids))
                row["reference_us"] = round(timed_us(lambda: theirs.search(q, rsp), 5, lambda: None), 1)
            out[f"{name}_Q{Q}"] = row
    return out


def nprobe_sweep(qb, idx, xq_d, xq_h, x, W, dev):
    """nprobe in {1, 4, 16, 64, 256}, recall@10 vs brute force beside device-resident QPS."""
    import torch
    xd = x.to(dev)
    gt = torch.cdist(xq_d, xd).topk(W["k"], largest=False).indices.cpu()
    del xd
    out = []
    for nprobe in (1, 4, 16, 64, 256):
        sp = qb.SearchParams()
        sp.k, sp.nprobe = W["k"], nprobe
        res = idx.search(xq_h, sp)
        for _ in range(3):
            idx._search_device(xq_d, sp)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(50):
            idx._search_device(xq_d, sp)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 50
        out.append({"nprobe": nprobe, "recall_at_10": round(recall_at_k(res.ids, gt), 4), "ms_per_step": round(ms, 4),
                    "qps": round(W["Q"] / (ms * 1e-3))})
    return out


def build_section(qb, x, W, dev, binfo, build_s):
    """k-means path: assign = 2*N*K*d flop per iteration against the tensor peak, update (counting sort +
    per-centroid sums) = N*d*4 + N*8 + K*d*4 bytes against the HBM peak. One Lloyd iteration on the bench's data."""
    import torch
    from quake_b200 import clustering, _lib
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    N, K, d = W["N"], W["nlist"], W["d"]
    xd = clustering.pad_rows(x, dev)
    cents = xd[torch.randperm(N, device=dev)[:K]].clone()
    filt = clustering.AssignFilter(dev)  # the precision policy of a training run (2xTF32 unless points get re-scanned)
    for _ in range(2):  # warm-up of all three stages (first launches load the kernels)
        a = clustering.assign_points(xd, d, cents, _lib.QK_METRIC_L2, filt=filt)
        filt.review()
        counts, offsets, order = clustering.partition_by_assignment(a, K)
        sums = clustering.centroid_sums(xd, d, order, offsets, K)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    torch.cuda.synchronize()
    e[0].record()
    a = clustering.assign_points(xd, d, cents, _lib.QK_METRIC_L2, filt=filt)
    e[1].record()
    counts, offsets, order = clustering.partition_by_assignment(a, K)
    e[2].record()
    sums = clustering.centroid_sums(xd, d, order, offsets, K)
    e[3].record()
    torch.cuda.synchronize()
    t_assign, t_sort, t_sum = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[2].elapsed_time(e[3])
    flop = 2.0 * N * K * d
    upd_bytes = N * d * 4 + N * 8 + K * d * 4
    tf_peak = float(peaks.get("bf16_tflops", 0)) or None
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    return {"build_s": round(build_s, 3), "train_time_us": getattr(binfo, "train_time_us", None),
            "assign": {"ms": round(t_assign, 3), "tflops": round(flop / (t_assign * 1e-3) / 1e12, 2),
                       "note": f"{filt.terms}xTF32 tensor-core filter + exact refine of the winner (k = 1); flop = 2*N*K*d",
                       "points_rescanned_exactly": int(filt.stats[0].item()),
                       "frac_of_bf16_dense_peak": round(flop / (t_assign * 1e-3) / 1e12 / tf_peak, 4) if tf_peak else None},
            "update": {"sort_ms": round(t_sort, 3), "sums_ms": round(t_sum, 3), "largest_list": int(counts.max()),
                       "note": "first Lloyd iteration (random initial centroids): list sizes are skewed, the per-centroid "
                               "sums are sequential in list order (reference order), so the longest list bounds the kernel",
                       "gbs": round(upd_bytes / ((t_sort + t_sum) * 1e-3) / 1e9, 1),
                       "frac_of_hbm_peak": round(upd_bytes / ((t_sort + t_sum) * 1e-3) / 1e9 / hbm, 4),
                       "algorithmic_bytes": upd_bytes}}


def sharded_section(qb, world, rank, dev, args):
    """The multi-GPU split the north star names: lists sharded over the ranks (partition id % world),
    centroids replicated, per-rank partial top-k exchanged over NVLink and merged. Workload: C4-shaped (d 96, mean list
    length 1526, nprobe 64, k 10, l2, Q 1024), 4M vectors per GPU (C4 itself is 12.5M per GPU on 8), built with the
    distributed build; strong scaling of one 1024-query batch. `equals_unsharded`: a small replicated index sharded in
    place must return exactly what the unsharded index returns, through the same collective."""
    import torch
    import torch.distributed as dist
    from quake_b200 import clustering
    from quake_b200.sharded import ShardedQuakeIndex
    per_gpu = int(os.environ.get("QK_BENCH_SHARD_N", 4_000_000))
    d, k, nprobe, Q = 96, 10, 64, 1024
    n_total = per_gpu * world
    nlist = max(64, int(round(n_total / 1526)))
    g = torch.Generator().manual_seed(1234 + rank)
    xl = torch.randn(per_gpu, d, generator=g)
    idl = torch.arange(rank * per_gpu, (rank + 1) * per_gpu, dtype=torch.int64)
    bp = qb.IndexBuildParams()
    bp.nlist, bp.metric, bp.niter = nlist, "l2", 3
    sh = ShardedQuakeIndex()
    torch.cuda.synchronize(); dist.barrier()
    log(f"sharded: distributed build of {n_total} x {d}, nlist {nlist}")
    t0 = time.perf_counter()
    sh.build(xl, idl, bp)
    torch.cuda.synchronize(); dist.barrier()
    build_s = time.perf_counter() - t0
    log(f"sharded: built in {build_s:.1f}s; searching")
    del xl
    g2 = torch.Generator().manual_seed(4321)
    q = torch.randn(Q, d, generator=g2)
    xq = clustering.pad_rows(q, dev)
    sp = qb.SearchParams()
    sp.k, sp.nprobe = k, nprobe
    for _ in range(5):
        sh.search_device(xq, sp)
    reps = 30
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); dist.barrier()
    e0.record()
    for _ in range(reps):
        sh.search_device(xq, sp)
    e1.record()
    torch.cuda.synchronize(); dist.barrier()
    total_ms = e0.elapsed_time(e1) / reps
    log(f"sharded: {total_ms:.3f} ms per batch; timing the pieces")
    # the pieces: local partial search alone, exchange + merge alone
    e0.record()
    for _ in range(reps):
        part = sh.search_partial(xq, sp)
    e1.record()
    torch.cuda.synchronize()
    partial_ms = e0.elapsed_time(e1) / reps
    dist.barrier()
    e0.record()
    for _ in range(reps):
        sh.exchange_and_merge(part[0], part[1], k)
    e1.record()
    torch.cuda.synchronize()
    exch_ms = e0.elapsed_time(e1) / reps
    t = torch.tensor([total_ms, partial_ms, exch_ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, partial_ms, exch_ms = [float(v) for v in t.cpu()]
    # every rank must hold the same answer
    ids, dd = sh.search_device(xq, sp)
    chk = torch.stack([ids.sum(), (ids * torch.arange(1, k + 1, device=dev)).sum()]).to(torch.float64)
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    ranks_agree = bool(torch.equal(lo, hi))
    log("sharded: equals_unsharded check")
    # equals_unsharded on a small replicated index
    torch.manual_seed(1234)
    xs = torch.randn(60000, d)
    bps = qb.IndexBuildParams()
    bps.nlist, bps.metric = 64, "l2"
    full = qb.QuakeIndex()
    full.build(xs, torch.arange(60000, dtype=torch.int64), bps)
    sps = qb.SearchParams()
    sps.k, sps.nprobe = k, 16
    want = full.search(q[:256], sps)
    sh2 = ShardedQuakeIndex()
    sh2.shard_from(full)
    got = sh2.search(q[:256], sps)
    eq = torch.tensor([int(torch.equal(got.ids, want.ids) and torch.equal(got.distances, want.distances))], device=dev)
    dist.all_reduce(eq, op=dist.ReduceOp.MIN)
    scan_bytes = None
    return {"workload": f"C4-shaped: {n_total} x {d} f32 randn ({per_gpu} per GPU), nlist={nlist}, Q={Q}, k={k}, l2, "
                        f"nprobe={nprobe}, lists sharded over {world} GPUs by partition id",
            "scaling": "strong", "build_s": round(build_s, 2), "ms_per_1024q": round(total_ms, 4),
            "qps": round(Q / (total_ms * 1e-3)), "partial_search_ms": round(partial_ms, 4),
            "exchange_and_merge_us": round(exch_ms * 1e3, 1), "exchange": sh.exchange_kind(),
            "ranks_agree": ranks_agree, "equals_unsharded": bool(int(eq.item()))}


def cpu_baseline(idx, ref, quake_ref, xq_h, W, ref_why):
    """The compiled reference (oracle/_ref) searching THE SAME index (saved by us in the reference's format,
    loaded by the reference) on the host cores; bounded sample."""
    cores = host_cores()
    if ref is not None:
        best = time_reference_search(ref, quake_ref, xq_h, W["k"], W["nprobe"], cores, budget_s=20.0)
        return {"value": best["value"], "unit": UNIT, "cores": cores, "kind": "reference",
                "sample": f"all {W['Q']} queries x {best['reps']} reps, same index (loaded from our save), "
                          f"{best['mode']}, num_threads={cores}; {reference_build_note()}"}
    from oracle import oracle as orc
    qn = 16
    t0 = time.perf_counter()
    orc.search_index_like(idx, xq_h[:qn], k=W["k"], nprobe=W["nprobe"])
    dt = time.perf_counter() - t0
    return {"value": qn / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{qn} queries, scalar oracle port (compiled reference unavailable: {ref_why}); includes list export"}


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
