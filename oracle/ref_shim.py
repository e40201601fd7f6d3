"""TEST INFRASTRUCTURE ONLY - loads the *unmodified* reference modules from /root/reference (build container) or from
the byte-for-byte copy ``oracle/make_ref.py`` leaves under ``oracle/_ref/`` (git-ignored; it travels to the GPU box
with the snapshot, where ``/root/reference`` does not exist).
It is used by ``scripts/make_golden.py`` to generate the committed fixtures under ``tests/golden/``, by the
container-only differential tests, and by ``bench.py`` to TIME the reference's own hot path on the box's host cores
(``--impl reference`` and ``cpu_baseline``, ``kind: "reference"``).  Nothing in the product package imports this file.

Recipe: the heavy package ``__init__`` files pull in scanpy /
plotnine / mudata, which are not installed.  We pre-register bare package modules whose
``__path__`` points into the reference tree, stub the four missing third-party modules,
and then import the hot-path modules themselves unchanged:

  liana/method/_pipe_utils/_get_mean_perms.py   (_get_means_perms, _calculate_pvals, ...)
  liana/method/sc/_liana_pipe.py                (liana_pipe, _run_method, _get_lr)
  liana/method/sc/_cellphonedb.py, _geometric_mean.py, _cellchat.py, _Method.py, _rank_aggregate.py
  liana/method/_pipe_utils/_aggregate.py, _pre.py, _common.py
  liana/resource/_reassemble_complexes.py, select_resource.py
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import pandas as pd

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_reference_root() -> str:
    """/root/reference in the build container; on the GPU box (where that tree does not exist) the unmodified copy that
    `oracle/make_ref.py` left under `oracle/_ref/` (git-ignored, shipped with the snapshot)."""
    env = os.environ.get("LIANA_REFERENCE_ROOT")
    if env:
        return env
    for cand in ("/root/reference", os.path.join(_HERE, "_ref")):
        if os.path.isdir(os.path.join(cand, "liana", "method", "sc")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _find_reference_root()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "liana", "method", "sc"))


class _DocstringProcessor:
    def __init__(self, *a, **k):
        self.params = {}

    def dedent(self, f):
        return f

    def __getattr__(self, name):  # any other decorator-like attribute is a pass-through
        def passthrough(*a, **k):
            if len(a) == 1 and callable(a[0]) and not k:
                return a[0]
            return lambda f: f
        return passthrough


def _scanpy_scale(adata, zero_center=True, max_value=None, copy=False, layer=None, obsm=None):
    """Restatement of scanpy.pp.scale (third-party dependency, pinned scanpy 1.9.8 `poetry.lock:3475`, not
    vendored under /root/reference) for the one call the reference makes: `sc.pp.scale(adata, copy=True).X`
    (`liana/method/sc/_liana_pipe.py:262`).  Published algorithm (scanpy/preprocessing/_simple.py
    `scale_sparse` -> `scale_array`, `_utils._get_mean_var`): densify, mean and E[x^2] per gene in float64
    (the square taken in the matrix dtype), unbiased variance, std 0 -> 1, then in-place `X -= mean; X /= std`
    on the float32 matrix."""
    import numpy as np
    from scipy import sparse
    assert zero_center and max_value is None and layer is None and obsm is None
    X = adata.X.toarray() if sparse.issparse(adata.X) else np.array(adata.X)
    mean = np.mean(X, axis=0, dtype=np.float64)
    mean_sq = np.multiply(X, X).mean(axis=0, dtype=np.float64)
    var = mean_sq - mean ** 2
    var *= X.shape[0] / (X.shape[0] - 1)
    std = np.sqrt(var)
    std[std == 0] = 1
    X -= mean
    X /= std
    out = adata.copy() if copy else adata
    out.X = X
    return out if copy else None


def _anndata_concat(adatas, axis=0, label=None, **kwargs):
    """Restatement of `anndata.concat` (third-party, pinned anndata 0.9.2, not vendored) for the one call the
    reference makes: `an.concat([xdata, ydata], axis=1, label='modality')` (`liana/utils/mdata_to_anndata.py:51`) -
    variables side by side, inner join on the cell names in the first object's order, `var[label]` = position."""
    from scipy import sparse
    from liana_py_b200.testing._anndata_lite import AnnDataLite
    assert axis == 1 and len(adatas) == 2
    x, y = adatas
    common = x.obs_names.intersection(y.obs_names, sort=False)
    xi, yi = x.obs_names.get_indexer(common), y.obs_names.get_indexer(common)
    X = sparse.hstack([sparse.csr_matrix(x.X)[xi], sparse.csr_matrix(y.X)[yi]], format="csr")
    var = pd.DataFrame(index=x.var_names.append(y.var_names))
    if label is not None:
        var[label] = pd.Categorical(["0"] * x.shape[1] + ["1"] * y.shape[1])
    return AnnDataLite(X=X, obs=pd.DataFrame(index=common), var=var)


def _install_pandas_compat():
    """Container drift, not a change of the reference: `_run_method` ends with `lr_res.drop([None], axis=1)`
    (`_liana_pipe.py:449-450`) to remove the column a method without a magnitude / specificity score wrote
    under the label None.  pandas 2.0.3 (pinned, `poetry.lock:2402`) keeps None as a label; pandas 3 turns it
    into NaN and then cannot find None.  Make that one call mean what it meant under the pinned pandas."""
    import pandas as pd
    if getattr(pd.DataFrame.drop, "_lrperm_compat", False):
        return
    orig = pd.DataFrame.drop

    def drop(self, labels=None, *args, **kwargs):
        if isinstance(labels, list) and len(labels) == 1 and labels[0] is None and kwargs.get("axis") == 1:
            return self.loc[:, ~self.columns.isna()]
        return orig(self, labels, *args, **kwargs)
    drop._lrperm_compat = True
    pd.DataFrame.drop = drop
    _install_chained_fillna_compat()


def _install_chained_fillna_compat():
    """Container drift again: with `return_all_lrs=True` the reference fills the left-merged flag columns through a
    CHAINED in-place call, `lr_res['lrs_to_keep'].fillna(value=False, inplace=True)` (`_reassemble_complexes.py:56-57`).
    Under the pinned pandas 2.0.3 `lr_res[col]` is a cached view, so the call updates the frame, and filling an object
    column of True / NaN silently downcasts it to bool.  pandas 3 (copy-on-write) makes the same line a no-op, and
    `~lr_res[lrs_to_keep]` (`_liana_pipe.py:389`) then fails.  Emulated here and only here: a column taken out of a
    frame remembers where it came from, and an in-place `fillna` on it writes the (downcast) result back."""
    import weakref
    import pandas as pd
    orig_getitem = pd.DataFrame.__getitem__
    orig_fillna = pd.Series.fillna

    def getitem(self, key):
        out = orig_getitem(self, key)
        if isinstance(key, str) and isinstance(out, pd.Series):
            object.__setattr__(out, "_lrperm_parent", (weakref.ref(self), key))
        return out

    def fillna(self, value=None, *args, inplace=False, **kwargs):
        if not inplace:
            return orig_fillna(self, value, *args, **kwargs)
        parent = getattr(self, "_lrperm_parent", None)
        frame = parent[0]() if parent is not None else None
        filled = orig_fillna(self, value, *args, **kwargs).infer_objects()
        if frame is not None:
            frame[parent[1]] = filled
            return None
        return orig_fillna(self, value, *args, inplace=True, **kwargs)

    pd.DataFrame.__getitem__ = getitem
    pd.Series.fillna = fillna


def _install_numpy_compat():
    """Container drift, not a change of the reference: `_lr_probability` calls `np.product`
    (`liana/method/sc/_cellchat.py:8`), an alias of `np.prod` that NumPy 2 removed (pinned: 1.24.4)."""
    import numpy as np
    if not hasattr(np, "product"):
        np.product = np.prod


_loaded = None


def load_reference():
    """Return a namespace with the reference's own functions (imported unmodified)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")

    repo_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if repo_root not in sys.path:
        sys.path.insert(0, repo_root)
    from liana_py_b200.testing._anndata_lite import AnnDataLite, MuDataLite

    lroot = os.path.join(REFERENCE_ROOT, "liana")

    def bare_pkg(name, path):
        m = types.ModuleType(name)
        m.__path__ = [path]
        m.__package__ = name
        sys.modules[name] = m
        return m

    bare_pkg("liana", lroot)
    bare_pkg("liana.method", os.path.join(lroot, "method"))
    bare_pkg("liana.method.sc", os.path.join(lroot, "method", "sc"))
    bare_pkg("liana.utils", os.path.join(lroot, "utils"))
    bare_pkg("liana.resource", os.path.join(lroot, "resource"))

    # third-party stubs
    an = types.ModuleType("anndata")
    an.AnnData = AnnDataLite
    an.concat = _anndata_concat
    sys.modules.setdefault("anndata", an)
    sc = types.ModuleType("scanpy")
    sc.AnnData = AnnDataLite
    sc.pp = types.SimpleNamespace(scale=_scanpy_scale)
    sys.modules.setdefault("scanpy", sc)
    mu = types.ModuleType("mudata")

    mu.MuData = MuDataLite      # used in isinstance checks (`_liana_pipe.py:126`)
    sys.modules.setdefault("mudata", mu)
    dr = types.ModuleType("docrep")
    dr.DocstringProcessor = _DocstringProcessor
    sys.modules.setdefault("docrep", dr)

    # names `_liana_pipe.py` imports from packages whose __init__ we skipped
    sys.modules["liana.utils"].mdata_to_anndata = \
        importlib.import_module("liana.utils.mdata_to_anndata").mdata_to_anndata
    rc = importlib.import_module("liana.resource._reassemble_complexes")
    sys.modules["liana.resource"].explode_complexes = rc.explode_complexes
    sys.modules["liana.resource"].filter_reassemble_complexes = rc.filter_reassemble_complexes

    _install_pandas_compat()
    _install_numpy_compat()
    ns = types.SimpleNamespace()
    ns.AnnData = AnnDataLite
    ns.reassemble = rc
    ns.select_resource = importlib.import_module("liana.resource.select_resource")
    ns.pipe_utils = importlib.import_module("liana.method._pipe_utils")
    ns.pre = importlib.import_module("liana.method._pipe_utils._pre")
    ns.common = importlib.import_module("liana.method._pipe_utils._common")
    ns.get_mean_perms = importlib.import_module("liana.method._pipe_utils._get_mean_perms")
    ns.aggregate = importlib.import_module("liana.method._pipe_utils._aggregate")
    ns.liana_pipe = importlib.import_module("liana.method.sc._liana_pipe")
    ns.Method = importlib.import_module("liana.method.sc._Method")
    ns.cellphonedb_mod = importlib.import_module("liana.method.sc._cellphonedb")
    ns.geometric_mean_mod = importlib.import_module("liana.method.sc._geometric_mean")
    ns.rank_aggregate_mod = importlib.import_module("liana.method.sc._rank_aggregate")
    ns.connectome_mod = importlib.import_module("liana.method.sc._connectome")
    ns.logfc_mod = importlib.import_module("liana.method.sc._logfc")
    ns.natmi_mod = importlib.import_module("liana.method.sc._natmi")
    ns.sca_mod = importlib.import_module("liana.method.sc._singlecellsignalr")
    ns.scseqcomm = importlib.import_module("liana.method.sc._scseqcomm").scseqcomm
    ns.cellchat_mod = importlib.import_module("liana.method.sc._cellchat")
    ns.cellchat = ns.cellchat_mod.cellchat
    # the default consensus of `liana/method/__init__.py:13-14` (that __init__ pulls in the spatial stack)
    ns.default_methods = [ns.cellphonedb_mod.cellphonedb, ns.connectome_mod.connectome, ns.logfc_mod.logfc,
                          ns.natmi_mod.natmi, ns.sca_mod.singlecellsignalr]
    ns.rank_aggregate = ns.rank_aggregate_mod.AggregateClass(ns.rank_aggregate_mod._rank_aggregate_meta,
                                                             methods=ns.default_methods)
    ns.cellphonedb = ns.cellphonedb_mod.cellphonedb
    ns.geometric_mean = ns.geometric_mean_mod.geometric_mean
    ns.constants = importlib.import_module("liana._constants")
    _loaded = ns
    return ns
