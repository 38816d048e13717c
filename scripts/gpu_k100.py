import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import quake_b200 as qb
from quake_b200 import index as qi
torch.manual_seed(1234)
n, d, nlist = 500000, 128, 819
x = torch.randn(n, d)
bp = qb.IndexBuildParams(); bp.nlist, bp.metric = nlist, "l2"
idx = qb.QuakeIndex(); idx.build(x, torch.arange(n), bp)
q = torch.randn(1024, d)
for k in (10, 50, 100, 200):
    sp = qb.SearchParams(); sp.k, sp.nprobe = k, 64
    os.environ["QK_SCAN_STATS"] = "1"; qi.GRAPHS_ENABLED = False
    idx.search(q, sp); torch.cuda.synchronize()
    st = qi.LAST_SCAN_STATS.cpu().tolist()
    os.environ.pop("QK_SCAN_STATS"); qi.GRAPHS_ENABLED = True
    idx.search(q, sp); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): idx.search(q, sp)
    torch.cuda.synchronize()
    print("k", k, "ms", (time.perf_counter() - t0) / 5 * 1e3, "rescanned", st[0], "max appended", st[1], "mean", ((st[3] << 32) | (st[2] & 0xffffffff)) / 1024, flush=True)
