"""pytest configuration: the `gpu` marker (tests that need a B200) and import paths."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a); run with -m gpu")
    # the native library is built in-tree and git-ignored: on a fresh checkout build it (nvcc cross-compiles
    # without a GPU) so that the ABI tests have something to load
    lib = os.path.join(ROOT, "quake_b200", "lib", "libquake_b200.so")
    if not os.path.exists(lib):
        try:
            import __graft_entry__ as entry
            entry.build_cuda()
        except Exception as e:  # pragma: no cover -- the ABI tests will then say what is missing
            print(f"[conftest] could not build {lib}: {e!r}")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def quake_ref():
    """The compiled, unmodified reference (oracle/_ref, built by oracle/build_ref.sh); tests that pin the
    oracle against it are skipped where it has not been built."""
    p = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.exists(os.path.join(p, "quake_ref", "_bindings.so")):
        pytest.skip("oracle/_ref not built")
    if p not in sys.path:
        sys.path.insert(0, p)
    try:
        import quake_ref as q
    except Exception as e:  # pragma: no cover
        pytest.skip(f"oracle/_ref not importable here: {e}")
    return q
