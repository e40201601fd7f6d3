import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


_GPU_OK = None


def _gpu_available() -> bool:
    """The CUDA library loads and an engine can be created on device 0 (checked once)."""
    global _GPU_OK
    if _GPU_OK is None:
        try:
            from liana_py_b200 import _engine
            _engine.load_library()
            eng = _engine.Engine(0)
            eng.close()
            _GPU_OK = True
        except Exception:
            _GPU_OK = False
    return _GPU_OK


def pytest_collection_modifyitems(config, items):
    """Tests marked `gpu` are SKIPPED (not errored) on a machine without the CUDA library or a CUDA device, unless they
    were asked for explicitly with `-m gpu` (the GPU box: a missing library must fail loudly there)."""
    if "gpu" in (config.getoption("-m") or ""):
        return
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if gpu_items and not _gpu_available():
        skip = pytest.mark.skip(reason="needs the CUDA library and a CUDA device")
        for it in gpu_items:
            it.add_marker(skip)


def load_golden(name):
    from scipy import sparse
    z = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    d = {k: z[k] for k in z.files}
    d["X"] = sparse.csr_matrix((d["data"], d["indices"], d["indptr"]), shape=tuple(d["shape"]))
    d["L"] = len(d["labels_cat"])
    return d


@pytest.fixture(scope="session", params=["toy700", "small300", "ragged90"])
def golden(request):
    return load_golden(request.param)


def load_golden_cellchat(name):
    from scipy import sparse
    z = np.load(os.path.join(GOLDEN, f"{name}_cellchat.npz"))
    d = {k: z[k] for k in z.files}
    d["X"] = sparse.csr_matrix((d["data"], d["indices"], d["indptr"]), shape=tuple(d["shape"]))
    d["L"] = int(d["labels"].max()) + 1
    d["name"] = name
    return d


@pytest.fixture(scope="session", params=["small300", "dense700", "ragged90"])
def golden_cc(request):
    return load_golden_cellchat(request.param)


@pytest.fixture(scope="session")
def engine():
    from liana_py_b200 import _engine
    eng = _engine.Engine(0)
    yield eng
    eng.close()
