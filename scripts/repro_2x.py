import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import quake_b200 as qb
from quake_b200 import index as qi
qi.GRAPHS_ENABLED = False
n, d, nlist = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
Q, nprobe = int(sys.argv[4]), int(sys.argv[5])
torch.manual_seed(1234)
x = torch.randn(n, d)
bp = qb.IndexBuildParams(); bp.nlist, bp.metric, bp.niter = nlist, "l2", 2
idx = qb.QuakeIndex(); idx.build(x, torch.arange(n, dtype=torch.int64), bp)
print("built; filter_terms", idx.store.filter_terms, flush=True)
torch.manual_seed(4321)
q = torch.randn(Q, d)
sp = qb.SearchParams(); sp.k, sp.nprobe = 10, nprobe
r = idx.search(q, sp)
torch.cuda.synchronize()
print("search ok", r.ids[0].tolist(), flush=True)
