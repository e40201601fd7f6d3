"""CPU: the C-ABI library loads and exports every symbol include/lrperm.h declares."""
import os
import re

import pytest

from liana_py_b200 import _engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "lrperm.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lrperm_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _engine.load_library()
    declared = _declared_symbols()
    assert len(declared) >= 12
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in include/lrperm.h but not exported"
    assert sorted(_engine.ABI_SYMBOLS) == declared
    assert lib.lrperm_abi_version() == 3


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_engine.EngineError, match="lrperm_create failed"):
        _engine.Engine(0)
