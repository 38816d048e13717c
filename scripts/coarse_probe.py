"""Coarse-scan (flat store, dense mode) timing and rescan statistics vs k. Debug aid."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import quake_b200 as qb
from quake_b200 import index as qi

torch.manual_seed(1234)
x = torch.randn(int(sys.argv[1]) if len(sys.argv) > 1 else 4096, 128) * 0.3
idx = qb.QuakeIndex()
idx.build(x, torch.arange(x.shape[0], dtype=torch.int64), qb.IndexBuildParams())
torch.manual_seed(4321)
q = torch.randn(1024, 128)
qd = qi.clustering.pad_rows(q, idx.store.device)
qi.GRAPHS_ENABLED = False
os.environ["QK_SCAN_STATS"] = "1"
for k in (16, 64, 100, 122, 128, 200, 240, 256, 300, 500):
    sp = qb.SearchParams(); sp.k = k
    for _ in range(2):
        idx._search_device(qd, sp)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        ids, dist, _ = idx._search_device(qd, sp)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5 * 1e6
    st = qi.LAST_SCAN_STATS.cpu().tolist()
    gt = torch.cdist(q, x).topk(k, largest=False).indices
    ok = bool(torch.equal(ids.cpu(), gt))
    print(f"k={k} us={dt:.0f} rescanned={st[0]} max_cand={st[1]} ids_equal_bruteforce={ok}", flush=True)
