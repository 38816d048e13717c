"""Second golden set from an AVX-512 build of the UNMODIFIED reference (the reference's own CMake uses -march=native,
which is AVX-512 on the survey host; oracle/_ref is built -march=x86-64-v3 so that it can run on any GPU box).

    ORACLE_MARCH=x86-64-v4 ORACLE_OUT=$PWD/oracle/_ref512 bash oracle/build_ref.sh
    python tests/golden/make_golden_avx512.py

GCC vectorises faiss's fvec_L2sqr / fvec_inner_product 16 lanes wide for AVX-512 instead of 8, so the summation
order -- and with it the last bits of every distance -- differs from the AVX2 build our kernels reproduce bit for
bit. This set pins what the north star asks for across that ISA boundary: ids identical, distances within 1e-4.
Outputs search_avx512.npz: the same searches as search.npz on the committed indexes index_l2 / index_ip, plus a
d = 128 index (index128_l2, built and saved here by the AVX-512 reference) with its queries and results.
"""
import os
import shutil
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref512"))
import quake_ref as quake  # noqa: E402  (the AVX-512 build)


def T(a):
    return a.numpy() if isinstance(a, torch.Tensor) else np.asarray(a)


S = np.load(os.path.join(HERE, "search.npz"))
res = {}
for m in ("l2", "ip"):
    idx = quake.QuakeIndex()
    idx.load(os.path.join(HERE, f"index_{m}"), 0)
    q = torch.from_numpy(S[f"{m}_q"])
    for tag in ("serial_small", "serial", "batched", "k100", "all"):
        nq, k, nprobe, batched = [int(v) for v in S[f"{m}_{tag}_cfg"]]
        sp = quake.SearchParams(); sp.k = k; sp.nprobe = nprobe; sp.batched_scan = bool(batched)
        r = idx.search(q[:nq], sp)
        res[f"{m}_{tag}_ids"], res[f"{m}_{tag}_dist"] = T(r.ids), T(r.distances)

# d = 128 (the headline dimension): 16 lanes x 8 blocks instead of 8 lanes x 16
torch.manual_seed(1234)
N, d, nlist = 2000, 128, 20
x = torch.randn(N, d)
ids = torch.arange(N, dtype=torch.int64) + 7
bp = quake.IndexBuildParams(); bp.nlist = nlist; bp.metric = "l2"; bp.niter = 5
idx = quake.QuakeIndex()
idx.build(x, ids, bp)
path = os.path.join(HERE, "index128_l2")
shutil.rmtree(path, ignore_errors=True)
idx.save(path)
torch.manual_seed(4321)
q = torch.randn(64, d)
res["d128_q"] = T(q)
for tag, (nq, k, nprobe) in {"serial_small": (8, 10, 5), "serial": (64, 10, 5), "all": (64, 10, nlist)}.items():
    sp = quake.SearchParams(); sp.k = k; sp.nprobe = nprobe
    r = idx.search(q[:nq], sp)
    res[f"d128_{tag}_ids"], res[f"d128_{tag}_dist"] = T(r.ids), T(r.distances)
    res[f"d128_{tag}_cfg"] = np.array([nq, k, nprobe, 0])
np.savez_compressed(os.path.join(HERE, "search_avx512.npz"), **res)
print("AVX-512 golden fixtures written")
