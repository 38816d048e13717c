"""The drop-in surface: every public name of the reference's pybind module (src/cpp/bindings/wrap.cpp:48-186,
here the compiled reference in oracle/_ref) exists on the quake_b200 mirror, parameter objects carry the same defaults,
and the harness wrapper (src/python/index_wrappers/quake.py) has the same methods with the same parameter names.
Host logic only -- no GPU; skipped where oracle/_ref (or /root/reference) is absent."""
import inspect
import os
import sys
import types

import pytest

CLASSES = ["IndexBuildParams", "SearchParams", "MaintenancePolicyParams", "SearchResult", "SearchTimingInfo",
           "BuildTimingInfo", "ModifyTimingInfo", "MaintenanceTimingInfo", "QuakeIndex"]


def _public(obj):
    return sorted(n for n in dir(obj) if not n.startswith("_"))


@pytest.mark.parametrize("cls", CLASSES)
def test_every_reference_name_exists(quake_ref, cls):
    import quake_b200 as qb
    ref_cls = getattr(quake_ref, cls)
    ours = getattr(qb, cls)()
    missing = [n for n in _public(ref_cls) if not hasattr(ours, n)]
    assert missing == [], f"{cls}: the reference exposes {missing}"


@pytest.mark.parametrize("cls", ["IndexBuildParams", "SearchParams", "MaintenancePolicyParams"])
def test_parameter_defaults_equal_the_reference(quake_ref, cls):
    import quake_b200 as qb
    ref, ours = getattr(quake_ref, cls)(), getattr(qb, cls)()
    for n in _public(ref):
        rv = getattr(ref, n)
        if callable(rv):
            continue
        ov = getattr(ours, n)
        if isinstance(rv, float):
            assert ov == pytest.approx(rv, rel=1e-6), (cls, n)
        else:
            assert ov == rv, (cls, n)


def test_harness_wrapper_signatures():
    ref_py = "/root/reference/src/python"
    if not os.path.isdir(ref_py):
        pytest.skip("reference tree not present")
    # the reference wrapper imports `quake` (its compiled bindings): a stand-in namespace is enough to read signatures
    saved = {k: sys.modules.get(k) for k in ("quake", "quake.index_wrappers", "quake.index_wrappers.wrapper",
                                             "quake.index_wrappers.quake")}
    pkg = types.ModuleType("quake")
    pkg.__path__ = [ref_py]
    for n in CLASSES:
        setattr(pkg, n, type(n, (), {}))
    sys.modules["quake"] = pkg
    old_flag = sys.dont_write_bytecode
    sys.dont_write_bytecode = True  # the reference tree is read-only
    try:
        import quake.index_wrappers.quake as refw
        from quake_b200 import workload as wl
        for name, fn in inspect.getmembers(refw.QuakeWrapper, predicate=inspect.isfunction):
            if name.startswith("_") and name != "__init__":
                continue
            assert hasattr(wl.QuakeWrapper, name), f"QuakeWrapper.{name} missing"
            want = list(inspect.signature(fn).parameters)
            got = list(inspect.signature(getattr(wl.QuakeWrapper, name)).parameters)
            extra_ok = {"_ignored"}
            assert [p for p in want if p not in got] == [], f"QuakeWrapper.{name}: parameters {want} vs {got}"
            assert all(p in want or p in extra_ok for p in got), f"QuakeWrapper.{name}: parameters {want} vs {got}"
    finally:
        sys.dont_write_bytecode = old_flag
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
