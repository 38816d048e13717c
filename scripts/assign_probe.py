"""Top-1 centroid assignment of add(): ours vs float64 brute force; gaps of the mismatches (debug aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import quake_b200 as qb
torch.manual_seed(1234)
n0, nlist = 400_000, 656
x = torch.randn(n0, 128)
bp = qb.IndexBuildParams(); bp.nlist, bp.metric, bp.niter = nlist, "l2", 3
idx = qb.QuakeIndex(); idx.build(x, torch.arange(n0, dtype=torch.int64), bp)
xa = torch.randn(40_000, 128)
sp = qb.SearchParams(); sp.k, sp.nprobe, sp.batched_scan = 1, nlist, True
res = idx.parent.search(xa, sp)
cents = idx.parent.get(torch.arange(nlist)).double().cuda()
d2 = torch.cdist(xa.double().cuda(), cents) ** 2
best = d2.argmin(1).cpu()
mism = torch.nonzero(res.ids[:, 0] != best).reshape(-1)
print("mismatches vs float64 argmin:", mism.numel())
for i in mism[:10].tolist():
    a, b = int(res.ids[i, 0]), int(best[i])
    print(i, a, b, float(d2[i, a]), float(d2[i, b]), "rel gap", float(abs(d2[i, a] - d2[i, b]) / d2[i, b]))
# remove: order semantics on this index
