"""Inputs of the call-signature variants shared by scripts/make_golden.py (which runs the reference on them in
the build container) and tests/test_dropin_gpu.py (which runs the drop-in on them on the GPU box)."""
import os

import numpy as np
import pandas as pd

from liana_py_b200.testing._synth import make_synthetic_adata

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


SUPP_COLUMNS = ["ligand_means_sums", "receptor_means_sums", "ligand_zscores", "receptor_zscores", "ligand_logfc",
                "receptor_logfc", "mat_mean", "mat_max", "ligand_trimean", "receptor_trimean", "ligand_cdf", "receptor_cdf"]


def variant_inputs():
    """small300 with the optional arguments of the call signature exercised (liana/tests/test_sc_methods.py,
    test_pre.py use the same knobs on pbmc68k): groupby_pairs, interactions=, resource=, .raw, layers."""
    adata = make_synthetic_adata(300, 400, 5, density=0.15, seed_data=11, seed_labels=12)
    cats = list(adata.obs["cluster"].cat.categories)
    pairs = pd.DataFrame({"source": [cats[0], cats[0], cats[2], cats[3]], "target": [cats[1], cats[0], cats[3], cats[0]]})
    res = pd.read_csv(os.path.join(ROOT, "liana_py_b200", "resource", "consensus.csv"))
    present = set(adata.var_names)
    ok = res[[all(g in present for g in (l + "_" + r).split("_")) for l, r in zip(res["ligand"], res["receptor"])]]
    interactions = list(zip(ok["ligand"].head(40), ok["receptor"].head(40)))
    custom = ok.iloc[40:140][["ligand", "receptor"]].reset_index(drop=True)
    # .raw holds log1p(2x) of X, a layer holds sqrt(X): three different matrices in one object
    X = adata.X.tocsr()
    raw = X.copy(); raw.data = np.log1p(2.0 * raw.data).astype(np.float32)
    lay = X.copy(); lay.data = np.sqrt(lay.data).astype(np.float32)
    adata.raw = type("Raw", (), {"X": raw, "var_names": adata.var_names, "var": adata.var})()
    adata.layers["sqrt"] = lay
    return adata, dict(
        pairs=dict(groupby_pairs=pairs, use_raw=False),
        interactions=dict(interactions=interactions, use_raw=False),
        resource=dict(resource=custom, use_raw=False),
        raw=dict(use_raw=True),
        layer=dict(layer="sqrt", use_raw=False),
        resname=dict(resource_name="CellPhoneDB", use_raw=False),     # another bundled resource, case-insensitive name
        # every statistic column another method would request, carried along as supplementary columns
        supp=dict(use_raw=False, supp_columns=SUPP_COLUMNS),
    )




def mdata_inputs():
    """The variant input split into two modalities of a MuData look-alike (liana/tests/test_sc_methods.py:167-190,
    test_rank_aggregate.py:72-92 call the methods on a MuData with `mdata_kwargs`): x takes its `sqrt` layer, y is
    transformed on the way in."""
    from liana_py_b200.testing._anndata_lite import MuDataLite
    adata, _ = variant_inputs()
    mdata = MuDataLite({"rna": adata[:, 0:200].copy(), "prot": adata[:, 200:400].copy()})
    return mdata, dict(x_mod="rna", y_mod="prot", x_layer="sqrt", y_transform=np.sqrt)


def nanlabel_inputs():
    """The variant input with every tenth cell (at random) left without a cluster label (NaN in `groupby`): the
    reference keeps such cells in the matrix - they are shuffled with the others, count in the whole-matrix
    statistics and in the "rest" of the log fold changes - but they belong to no cluster."""
    adata, _ = variant_inputs()
    lab = adata.obs["cluster"].astype(object).copy()
    lab[np.random.default_rng(3).random(len(lab)) < 0.1] = np.nan
    adata.obs["cluster"] = pd.Categorical(lab)
    return adata


def signed_inputs():
    """The variant input shifted so that a third of the stored values are negative (scaled / centred data) and with
    INTEGER cluster names (1, 4, 7, ...): `source` / `target` come back as int64 columns in the reference."""
    adata, _ = variant_inputs()
    X = adata.X.copy()
    X.data = (X.data - 1.0).astype(np.float32)
    adata.X = X
    adata.obs["cluster"] = pd.Categorical(adata.obs["cluster"].cat.codes.astype(int) * 3 + 1)
    return adata


def sample_inputs():
    """The variant input with a plain-string `sample` column (two unequal samples): `by_sample` converts it to a
    categorical, runs every sample on its own rows and stacks the frames in category order."""
    adata, _ = variant_inputs()
    adata.obs["sample"] = np.where(np.arange(adata.shape[0]) % 3 == 0, "s2", "s1")
    return adata


def ties_inputs():
    """The variant input with TIED subunits: for every complex of the consensus resource whose subunits are all
    present, the first subunit's expression is copied into the others (equal cluster means and proportions in every
    cluster: which subunit represents the complex is decided by the reference's tie rule alone); for every second such
    complex the copy is then zeroed in the cells of the first cluster, so the tie holds in some clusters only."""
    adata, _ = variant_inputs()
    res = pd.read_csv(os.path.join(ROOT, "liana_py_b200", "resource", "consensus.csv"))
    names = list(adata.var_names)
    pos = {g: i for i, g in enumerate(names)}
    X = adata.X.tolil(copy=True)
    first = np.flatnonzero(np.asarray(adata.obs["cluster"].cat.codes) == 0)
    seen, k = set(), 0
    for entry in list(res["ligand"]) + list(res["receptor"]):
        subs = entry.split("_")
        if len(subs) < 2 or entry in seen or not all(g in pos for g in subs):
            continue
        seen.add(entry)
        for g in subs[1:]:
            X[:, pos[g]] = X[:, pos[subs[0]]]
            if k % 2:
                X[first, pos[g]] = 0
        k += 1
    adata.X = X.tocsr().astype(np.float32)
    adata.X.eliminate_zeros()
    return adata, k
