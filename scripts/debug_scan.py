"""GPU debugging aid: run one scan parity case verbosely."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tests.test_gpu_scan as t
from oracle import oracle as orc
from quake_b200 import index as qidx, clustering, _lib

d = int(sys.argv[1]); metric = sys.argv[2]
k = int(sys.argv[3]) if len(sys.argv) > 3 else 10
mode = sys.argv[4] if len(sys.argv) > 4 else "ragged"
if mode == "ragged":
    g = np.random.default_rng(d)
    sizes = g.integers(0, 700, size=64); sizes[3] = 0; sizes[5] = 1
    st, lists = t._make_store(sizes, d, seed=d)
    Q, nprobe, seed = 200, 8, 1
else:
    sizes = np.full(32, 300)
    st, lists = t._make_store(sizes, d, seed=k)
    Q, nprobe, seed = 70, 6, 2
gg = torch.Generator().manual_seed(seed)
q = torch.randn(Q, d, generator=gg)
L = len(lists)
probe = torch.stack([torch.randperm(L, generator=gg)[:nprobe] for _ in range(Q)]).to(torch.int64)
m = _lib.QK_METRIC_INNER_PRODUCT if metric == "ip" else _lib.QK_METRIC_L2
xq = clustering.pad_rows(q, st.device)
ids, dist, rows = qidx.scan_partitions(st, xq, probe.to(torch.int32).to(st.device), k, m, want_rows=True)
torch.cuda.synchronize()
oi, od, _ = orc.serial_scan(q.numpy(), lists, probe.numpy(), k, metric)
ids, dist = ids.cpu().numpy(), dist.cpu().numpy()
bad = np.nonzero((ids != oi).any(axis=1))[0]
print("bad queries", bad)
np.set_printoptions(linewidth=200)
for b in bad[:3]:
    col = np.nonzero(ids[b] != oi[b])[0]
    print("q", b, "first differing rank", col[:10], "missing from gpu:", sorted(set(oi[b]) - set(ids[b]))[:10], "extra in gpu:", sorted(set(ids[b]) - set(oi[b]))[:10])
for b in bad[:5]:
    print("q", b, "probe", probe[b].tolist(), "sizes", [int(sizes[p]) for p in probe[b]])
    print(" gpu ids ", ids[b]); print(" ref ids ", oi[b])
    print(" gpu dist", dist[b]); print(" ref dist", od[b])
