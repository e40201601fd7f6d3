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


def test_product_never_imports_the_oracle():
    """`oracle/` and the tests' engine stand-in are checkers: nothing under the package may import them, and the
    engine has no CPU path (creating one without the CUDA library / a device raises)."""
    import ast
    import pathlib
    root = pathlib.Path(__file__).resolve().parent.parent / "liana_py_b200"
    for path in root.rglob("*.py"):
        tree = ast.parse(path.read_text())
        for node in ast.walk(tree):
            names = []
            if isinstance(node, ast.Import):
                names = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                names = [node.module or ""]
            for n in names:
                assert not (n == "oracle" or n.startswith("oracle.") or n.startswith("_oracle_engine")), (str(path), n)
    import torch
    if not torch.cuda.is_available():
        import pytest
        from liana_py_b200 import _engine
        with pytest.raises(_engine.EngineError):
            _engine.Engine(0)
