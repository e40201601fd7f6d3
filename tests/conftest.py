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
