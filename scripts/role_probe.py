"""Per-role wait cycles of scan_mma_kernel (debug build, -DQK_STAGE_DEBUG): where every warp role of the pipeline
spends its time, averaged over the 148 CTAs, for the partition scan and the coarse (flat) scan of one C2 step.
usage: QK_LIB_PATH=quake_b200/lib/libquake_b200_dbg.so python scripts/role_probe.py [N nlist Q nprobe]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import quake_b200 as qb
from quake_b200 import index as qi, _lib
qi.GRAPHS_ENABLED = False
a = [int(v) for v in sys.argv[1:]] + [None] * 4
n, nlist, Q, nprobe = a[0] or 1000000, a[1] or 4096, a[2] or 1024, a[3] or 64
torch.manual_seed(1234)
x = torch.randn(n, 128)
bp = qb.IndexBuildParams(); bp.nlist, bp.metric, bp.niter = nlist, "l2", 5
idx = qb.QuakeIndex(); idx.build(x, torch.arange(n, dtype=torch.int64), bp)
torch.manual_seed(4321)
xq = torch.randn(Q, 128).cuda()
sp = qb.SearchParams(); sp.k, sp.nprobe = 10, nprobe
lib = _lib.load()
fn = lib.qk_debug_times
fn.restype = C.c_int
fn.argtypes = [C.c_void_p, C.c_int]
for _ in range(4):
    idx._search_device(xq, sp)
torch.cuda.synchronize()
fn(None, 1)
REPS = 5
for _ in range(REPS):
    idx._search_device(xq, sp)
torch.cuda.synchronize()
buf = np.zeros(2 * 148 * 64, dtype=np.uint64)
fn(buf.ctypes.data_as(C.c_void_p), 0)
t = buf.reshape(2, 148, 64).astype(np.float64) / REPS / 1965.0  # us at 1965 MHz
names = {0: ("producer", ["total", "wait a_empty", "wait i_empty", "tma issue", "descriptor (incl. i_empty)", "metadata rotate"]),
         8: ("mma issuer 0", ["total", "wait i_full", "wait b_ready", "wait d_empty", "wait alo_full"]),
         16: ("mma issuer 1", ["total", "wait i_full", "wait b_ready", "wait d_empty", "wait alo_full"]),
         24: ("split warp 2", ["total", "wait i_full", "wait b_empty", "wait a_full"]),
         32: ("epilogue warp 6 (group 0)", ["total", "wait i_full", "wait d_full", "tile work (ld+score+append)"]),
         40: ("epilogue warp 10 (group 1)", ["total", "wait i_full", "wait d_full", "tile work (ld+score+append)"])}
for kind, label in ((0, "partition scan"), (1, "coarse scan (flat, dense)")):
    print(f"== {label}: mean over CTAs, us per launch (min..max)")
    for base, (role, cols) in names.items():
        parts = []
        for i, c in enumerate(cols):
            v = t[kind, :, base + i]
            parts.append(f"{c} {v.mean():.1f} ({v.min():.1f}..{v.max():.1f})")
        print(f"  {role:28s} " + "; ".join(parts))
