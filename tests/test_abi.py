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
    assert lib.lrperm_abi_version() == 5


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


def test_ctypes_signatures_match_the_header():
    """Every prototype of include/lrperm.h against the `argtypes` the ctypes binding declares: same number of
    parameters, pointers where the header has pointers, 64-bit integers where it has int64_t / uint64_t, doubles and
    floats where it has them - a binding that drifts from the header corrupts memory silently."""
    import ctypes as C
    import re
    from liana_py_b200 import _engine
    lib = _engine.load_library()
    header = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "lrperm.h")).read()
    header = re.sub(r"/\*.*?\*/", " ", header, flags=re.S)
    protos = re.findall(r"\b(?:int|void|const\s+char\s*\*)\s*(lrperm_\w+)\s*\(([^;{}]*?)\)\s*;", header)
    assert {p[0] for p in protos} == set(_engine.ABI_SYMBOLS)
    checked = 0
    for name, params in protos:
        fn = getattr(lib, name)
        plist = [] if params.strip() in ("", "void") else [p.strip() for p in params.split(",")]
        if fn.argtypes is None:
            assert name in ("lrperm_abi_version",) or not plist, f"{name}: no argtypes declared"
            continue
        assert len(fn.argtypes) == len(plist), (name, len(fn.argtypes), plist)
        for at, p in zip(fn.argtypes, plist):
            if "*" in p:
                assert at in (C.c_void_p, C.c_char_p) or issubclass(at, C._Pointer), (name, p, at)
            elif re.search(r"\b(u?int64_t)\b", p):
                assert C.sizeof(at) == 8 and at not in (C.c_double, C.c_void_p), (name, p, at)
            elif re.search(r"\bdouble\b", p):
                assert at is C.c_double, (name, p, at)
            elif re.search(r"\bfloat\b", p):
                assert at is C.c_float, (name, p, at)
            else:                                   # int / int32_t / enums
                assert C.sizeof(at) == 4 and at not in (C.c_float,), (name, p, at)
        checked += 1
    assert checked >= 25
    # return types: status codes are int, the error text a C string, destroy returns nothing
    rets = {v: k for k, v in re.findall(r"\b(int|void|const\s+char\s*\*)\s*(lrperm_\w+)\s*\(", header)}
    for name, ret in rets.items():
        rt = getattr(lib, name).restype
        if ret == "void":
            assert rt is None, name
        elif ret.startswith("const"):
            assert rt is C.c_char_p, name
        else:
            assert rt is C.c_int, name
