"""APS timing on a C3-shaped index (ip, k 100, recall 0.9, n-bar 610); QK_APS_TRACE=1 prints the rounds."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import quake_b200 as qb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
nlist = 16384 if n == 10_000_000 else n // 610  # C3: 10M vectors, nlist 16384
torch.manual_seed(1234)
x = torch.randn(n, 128); x /= x.norm(dim=1, keepdim=True)
torch.manual_seed(4321)
q = torch.randn(1024, 128); q /= q.norm(dim=1, keepdim=True)
bp = qb.IndexBuildParams(); bp.nlist, bp.metric, bp.niter = nlist, "ip", 2
idx = qb.QuakeIndex()
t0 = time.perf_counter(); idx.build(x, torch.arange(n, dtype=torch.int64), bp); torch.cuda.synchronize()
print(f"build {time.perf_counter() - t0:.1f}s nlist {nlist}", flush=True)
sp = qb.SearchParams(); sp.k, sp.recall_target, sp.initial_search_fraction = 100, 0.9, 0.02
idx.search(q[:64], sp)
for mode in ("1", "0"):
    os.environ["QK_APS_COLLECT"] = mode
    idx.search(q, sp)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = idx.search(q, sp)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"collect={mode}: {dt * 1e3:.2f} ms / 1024 q, mean partitions scanned {float(idx.last_partitions_scanned.float().mean()):.1f}", flush=True)
    if mode == "1":
        ref_ids = r.ids.clone()
    else:
        print("ids equal between modes:", bool(torch.equal(ref_ids, r.ids)))

# where the time goes outside the rounds: the candidate search in the parent (k = 2 % of the centroids)
from quake_b200 import clustering
os.environ["QK_APS_COLLECT"] = "1"
xq = clustering.pad_rows(q, idx.store.device)
psp = qb.SearchParams(); psp.batched_scan = True; psp.recall_target = 0.9
psp.k = max(int(idx.nlist() * 0.02), 1)
for _ in range(2):
    idx.parent._search_device(xq, psp, want_rows=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    idx.parent._search_device(xq, psp, want_rows=True)
torch.cuda.synchronize()
print(f"candidate search in the parent (k = {psp.k}): {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms", flush=True)
for _ in range(2):
    idx._search_device(xq, sp)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    idx._search_device(xq, sp)
torch.cuda.synchronize()
print(f"whole APS search, device-resident queries: {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms", flush=True)
