"""k-means assign speed vs K (debug aid): time per call, TFLOP/s; QK_PROBE_NCU=1 brackets ONE call for an ncu list."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from quake_b200 import clustering, _lib
dev = torch.device("cuda", 0)
torch.manual_seed(0)
n, d = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000, int(os.environ.get('D', '128'))
x = clustering.pad_rows(torch.randn(n, d), dev)
for K in [int(v) for v in (sys.argv[2:] or ["4096", "16384"])]:
    c = x[torch.randperm(n, device=dev)[:K]].clone()
    filt = clustering.AssignFilter(dev)
    for _ in range(2):
        a = clustering.assign_points(x, d, c, _lib.QK_METRIC_L2, filt=filt)
    torch.cuda.synchronize()
    if os.environ.get("QK_PROBE_NCU") == "1":
        torch.cuda.cudart().cudaProfilerStart()
    t0 = time.perf_counter()
    a = clustering.assign_points(x, d, c, _lib.QK_METRIC_L2, filt=filt)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if os.environ.get("QK_PROBE_NCU") == "1":
        torch.cuda.cudart().cudaProfilerStop()
    print(f"K={K} n={n} d={d} terms={filt.terms}: {dt * 1e3:.1f} ms, {2.0 * n * K * d / dt / 1e12:.1f} TFLOP/s, "
          f"re-scanned {int(filt.stats[0])} of {filt.points} points, max candidates {int(filt.stats[1])}", flush=True)
