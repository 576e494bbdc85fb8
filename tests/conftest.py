import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # a fresh checkout has no built artefacts (*.so are git-ignored): build them once, exactly as __graft_entry__.build() does
    lib = os.path.join(ROOT, "too-many-cells_b200", "libtmc.so")
    if not os.path.exists(lib):
        import subprocess
        subprocess.run(["make", "-C", os.path.join(ROOT, "too-many-cells_b200", "csrc")], check=True, stdout=subprocess.DEVNULL)


def load_golden(name):
    import numpy as np
    import scipy.sparse as sp
    z = np.load(os.path.join(GOLDEN, f"synth_{name}.npz"))
    n, g = (int(x) for x in z["shape"])
    a = sp.csr_matrix((z["counts"].astype(np.float64), z["indices"], z["indptr"]), shape=(n, g))
    return a, z


@pytest.fixture(scope="session")
def golden_tiny():
    return load_golden("tiny")


@pytest.fixture(scope="session")
def golden_small():
    return load_golden("small")


def pytest_collection_modifyitems(config, items):
    """`@pytest.mark.gpu` tests need a CUDA device: skip them (instead of failing) where there is none, so a plain `pytest`
    works on a CPU box; on the B200 box they run and load libtmc.so."""
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if not gpu_items:
        return
    try:
        from too_many_cells_b200 import _lib
        have = _lib.load().tmc_device_count() >= 1
    except Exception:
        have = False
    if not have:
        skip = pytest.mark.skip(reason="no CUDA device (libtmc has no CPU fallback)")
        for it in gpu_items:
            it.add_marker(skip)
