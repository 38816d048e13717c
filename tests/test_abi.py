"""The C-ABI shared library loads without a GPU and exports every symbol include/quake_b200.h declares
(no compute calls here)."""
import ctypes
import os
import re

import pytest

from tests.conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "quake_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(qk_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from quake_b200 import _lib

    names = _declared()
    assert len(names) >= 15
    lib = _lib.load()
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/quake_b200.h but not exported"
        assert n in _lib.PROTOTYPES, f"{n} has no ctypes prototype"
    assert sorted(_lib.PROTOTYPES) == names


def test_version_and_error_strings():
    from quake_b200 import _lib

    lib = _lib.load()
    assert b"sm_100a" in lib.qk_version()
    assert isinstance(lib.qk_last_error(), bytes)


def test_store_struct_layout_matches_header():
    """ctypes mirror of struct qk_store: field order/types as in the header."""
    from quake_b200._lib import QkStore

    src = open(os.path.join(ROOT, "include", "quake_b200.h")).read()
    body = src[src.index("typedef struct qk_store {"):src.index("} qk_store_t;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"(\w+)\s*;", body)
    assert fields == [f[0] for f in QkStore._fields_]
    assert ctypes.sizeof(QkStore) == 120


def test_no_gpu_means_loud_failure():
    """Without a compute-capability-10.x device every compute path raises; nothing falls back to the CPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import quake_b200 as qb

    idx = qb.QuakeIndex()
    bp = qb.IndexBuildParams()
    with pytest.raises(RuntimeError):
        idx.build(torch.randn(100, 8), torch.arange(100), bp)


def test_host_helpers_rand_perm():
    """qk_host_rand_perm_prefix == faiss::rand_perm (faiss/utils/random.cpp:153-163): mt19937 Fisher-Yates."""
    import numpy as np
    from quake_b200 import clustering

    n, seed = 1000, 1234
    # independent restatement with numpy's MT19937 raw 32-bit outputs (same generator as std::mt19937)
    bg = np.random.MT19937()
    bg._legacy_seeding(seed)
    raw = bg.random_raw
    perm = np.arange(n)
    for i in range(n - 1):
        i2 = i + int(raw()) % (n - i)
        perm[i], perm[i2] = perm[i2], perm[i]
    got = clustering.rand_perm_prefix(n, seed, 300)
    assert got.tolist() == perm[:300].tolist()
    assert sorted(clustering.rand_perm_prefix(50, 7, 50).tolist()) == list(range(50))


def test_host_split_clusters_rule():
    """qk_host_split_clusters on a hand-made case: one empty cluster takes a perturbed copy of the roulette's pick and
    half of its weight (Clustering.cpp:218-247)."""
    import ctypes as C
    import numpy as np
    from quake_b200 import _lib
    lib = _lib.load()
    d, k, n = 4, 3, 100
    hassign = np.array([60.0, 0.0, 40.0], dtype=np.float32)
    cents = np.array([[1, 2, 3, 4], [9, 9, 9, 9], [5, 6, 7, 8]], dtype=np.float32)
    nsplit = C.c_int64(0)
    rc = lib.qk_host_split_clusters(d, k, n, hassign.ctypes.data_as(C.POINTER(C.c_float)),
                                    cents.ctypes.data_as(C.POINTER(C.c_float)), d, C.byref(nsplit))
    assert rc == 0 and nsplit.value == 1
    # std::mt19937(1234): the first draw decides whether cluster 0 (p = 59/97) is taken; whichever cluster cj was taken,
    # the new centroid is cj * (1 +- 1/1024) on alternating dimensions and the weights are halved
    eps = 1.0 / 1024
    picked = 0 if abs(hassign[0] - 30.0) < 1e-6 else 2
    src = np.array([[1, 2, 3, 4], [9, 9, 9, 9], [5, 6, 7, 8]], dtype=np.float32)[picked]
    sign = np.array([1, -1, 1, -1], dtype=np.float32)
    assert np.allclose(cents[1], src * (1 + sign * eps), rtol=1e-6)
    assert np.allclose(cents[picked], src * (1 - sign * eps), rtol=1e-6)
    assert hassign[1] == [60.0, 0, 40.0][picked] / 2 and hassign[picked] == [60.0, 0, 40.0][picked] / 2
