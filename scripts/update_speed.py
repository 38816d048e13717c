"""k-means update kernels, warm (debug aid): counting sort by assignment + per-centroid sums, ms and GB/s."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from quake_b200 import clustering, _lib
dev = torch.device("cuda", 0)
torch.manual_seed(0)
n, d, K = 1_000_000, 128, 4096
x = clustering.pad_rows(torch.randn(n, d), dev)
c = x[torch.randperm(n, device=dev)[:K]].clone()
a = clustering.assign_points(x, d, c, _lib.QK_METRIC_L2)
for rep in range(3):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    torch.cuda.synchronize()
    e[0].record()
    counts, offsets, order = clustering.partition_by_assignment(a, K)
    e[1].record()
    sums = clustering.centroid_sums(x, d, order, offsets, K)
    e[2].record()
    torch.cuda.synchronize()
    ts, tu = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
    b = n * d * 4 + n * 8 + K * d * 4
    print(f"rep {rep}: sort {ts:.3f} ms, sums {tu:.3f} ms ({n * d * 4 / tu / 1e6:.0f} GB/s), update {b / (ts + tu) / 1e6:.0f} GB/s; max list {int(counts.max())}", flush=True)
