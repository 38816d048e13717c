"""Eager vs graph-replayed step time at several nprobe values, with and without the hit window (debug aid)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import quake_b200 as qb
from quake_b200 import index as qi
torch.manual_seed(1234)
x = torch.randn(1_000_000, 128)
bp = qb.IndexBuildParams(); bp.nlist, bp.metric = 4096, "l2"
idx = qb.QuakeIndex(); idx.build(x, torch.arange(x.shape[0], dtype=torch.int64), bp)
torch.manual_seed(4321)
qd = qi.clustering.pad_rows(torch.randn(1024, 128), idx.store.device)
def loop(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6
for nprobe in (64, 256):
    sp = qb.SearchParams(); sp.k, sp.nprobe = 10, nprobe
    for hits in (True, False):
        idx.maintenance_policy_params = qb.MaintenancePolicyParams() if hits else None
        qi.GRAPHS_ENABLED = False
        e = loop(lambda: idx._search_device(qd, sp))
        qi.GRAPHS_ENABLED = True
        g = loop(lambda: idx._search_device(qd, sp))
        plan = idx._plan(1024, sp)
        r = loop(lambda: plan.graph.replay())
        print(f"nprobe={nprobe} hits={hits} eager_us={e:.0f} graph_us={g:.0f} bare_replay_us={r:.0f}", flush=True)
