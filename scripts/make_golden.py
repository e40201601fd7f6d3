"""Container-only: generate tests/golden/*.npz by running the UNMODIFIED reference code
(/root/reference, loaded through oracle/ref_shim.py) on small synthetic inputs.

For each case we record the inputs the reference's hot path actually saw (the CSR matrix after
prep_check_adata + gene subsetting, label codes), the permutation cube it produced
(`_get_means_perms`, liana/method/_pipe_utils/_get_mean_perms.py:9-57), the index arrays of
`_get_mat_idx` (:103-113) and the final li.mt.cellphonedb / li.mt.geometric_mean frames.

Run:  python scripts/make_golden.py          (needs /root/reference; not runnable on the GPU box)
"""
from __future__ import annotations

import os
import sys

import numpy as np
import pandas as pd
from scipy import sparse

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.ref_shim import load_reference  # noqa: E402
from liana_py_b200.testing._synth import make_synthetic_adata  # noqa: E402
from liana_py_b200.testing._anndata_lite import AnnDataLite  # noqa: E402
sys.path.insert(0, os.path.join(ROOT, "tests"))
from variants import mdata_inputs, nanlabel_inputs, sample_inputs, signed_inputs, ties_inputs, variant_inputs  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def capture_run(ref, method, adata, **kw):
    """Run one reference method, capturing what the hot path saw and produced."""
    cap = {}
    pipe = ref.liana_pipe
    orig_perms, orig_idx = pipe._get_means_perms, pipe._get_mat_idx

    def wrap_perms(adata, n_perms, seed, agg_fun, norm_factor, n_jobs, verbose):
        cap["X"] = adata.X.copy().tocsr()
        cap["var_names"] = np.array(adata.var_names)
        cap["labels_cat"] = np.array(adata.obs['@label'].cat.categories)
        cap["labels"] = np.asarray(adata.obs['@label'].cat.codes, dtype=np.int32)
        cube = orig_perms(adata=adata, n_perms=n_perms, seed=seed, agg_fun=agg_fun, norm_factor=norm_factor,
                          n_jobs=n_jobs, verbose=verbose)
        cap["cube"] = cube
        return cube

    def wrap_idx(adata, lr_res):
        out = orig_idx(adata, lr_res)
        cap["lr_res_in"] = lr_res.copy()
        cap["idx"] = [np.asarray(o, dtype=np.int32) for o in out]
        return out

    pipe._get_means_perms, pipe._get_mat_idx = wrap_perms, wrap_idx
    try:
        res = method(adata, inplace=False, **kw)
    finally:
        pipe._get_means_perms, pipe._get_mat_idx = orig_perms, orig_idx
    cap["res"] = res
    return cap


def save_case(name, adata, groupby, n_perms, seed, expr_prop, ref, keep_cube_perms=None):
    kw = dict(groupby=groupby, use_raw=False, n_perms=n_perms, seed=seed, expr_prop=expr_prop, verbose=False)
    cp = capture_run(ref, ref.cellphonedb, adata.copy(), **kw)
    gm = capture_run(ref, ref.geometric_mean, adata.copy(), **kw)
    assert np.array_equal(cp["cube"], gm["cube"]), "both methods must see the same cube"
    X = cp["X"]
    X.sort_indices()
    N = X.shape[0]
    rng = np.random.default_rng(seed)
    perm_idx = np.stack([rng.permutation(np.arange(N)) for _ in range(n_perms)])
    cube = cp["cube"]
    assert np.array_equal(cube.astype(np.float32).astype(np.float64), cube), "cube holds float32 values"
    kc = n_perms if keep_cube_perms is None else min(keep_cube_perms, n_perms)
    lr_in = cp["lr_res_in"]
    lig, rec, src, tgt = cp["idx"]
    full = adata.X.tocsr()
    full.sort_indices()
    arrays = dict(
        # the full input (before the pipeline touches it) for drop-in tests
        in_indptr=full.indptr.astype(np.int64), in_indices=full.indices.astype(np.int32),
        in_data=full.data.astype(np.float32), in_shape=np.array(full.shape, dtype=np.int64),
        in_var_names=np.array(adata.var_names, dtype=str), in_labels=np.array([str(v) for v in adata.obs[groupby]], dtype=str),
        # what the hot path saw
        indptr=X.indptr.astype(np.int64), indices=X.indices.astype(np.int32), data=X.data.astype(np.float32),
        shape=np.array(X.shape, dtype=np.int64), var_names=cp["var_names"].astype(str),
        labels=cp["labels"], labels_cat=cp["labels_cat"].astype(str),
        perm_idx=perm_idx.astype(np.int32), cube=cube[:kc].astype(np.float32),
        cube_cluster_sums=cube.sum(axis=(0, 2)),     # like liana/tests/test_get_mean_perms.py:27-35
        ligand_idx=lig, receptor_idx=rec, source_idx=src, target_idx=tgt,
        ligand_means=lr_in["ligand_means"].values.astype(np.float32),
        receptor_means=lr_in["receptor_means"].values.astype(np.float32),
        n_perms=np.int64(n_perms), seed=np.int64(seed), expr_prop=np.float64(expr_prop),
    )
    # final frames, aligned to the order of lr_res_in (the rows the score function saw)
    key = ["source", "target", "ligand_complex", "receptor_complex"]
    for tag, cap_, mag, spec in (("cpdb", cp, "lr_means", "cellphone_pvals"), ("gmean", gm, "lr_gmeans", "gmean_pvals")):
        res = cap_["res"]
        merged = lr_in[key].merge(res[key + [mag, spec]], on=key, how="left")
        assert len(merged) == len(lr_in) and not merged[spec].isna().any()
        arrays[f"{tag}_magnitude"] = merged[mag].values.astype(np.float32)
        arrays[f"{tag}_pvals"] = merged[spec].values.astype(np.float64)
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **arrays)
    cp["res"].to_csv(os.path.join(OUT, f"{name}_cellphonedb.csv.gz"), index=False)
    gm["res"].to_csv(os.path.join(OUT, f"{name}_geometric_mean.csv.gz"), index=False)
    print(f"{name}: N={N} G={X.shape[1]} L={len(cp['labels_cat'])} P={n_perms} T={len(lig)} "
          f"mean cpdb pval={arrays['cpdb_pvals'].mean():.6f} mean gmean pval={arrays['gmean_pvals'].mean():.6f}")


def save_cellchat_case(name, adata, groupby, n_perms, seed, expr_prop, ref):
    """li.mt.cellchat of the reference: what `_get_means_perms` saw (the matrix BEFORE its in-place division,
    `_get_mean_perms.py:43-44`), the float64 trimean cube, the rows / indices / observed trimeans that reached
    `_cellchat_score`, and the final frame."""
    kw = dict(groupby=groupby, use_raw=False, n_perms=n_perms, seed=seed, expr_prop=expr_prop, verbose=False)
    cc = capture_run(ref, ref.cellchat, adata.copy(), **kw)
    X = cc["X"]
    X.sort_indices()
    lr_in = cc["lr_res_in"]
    lig, rec, src, tgt = cc["idx"]
    res = cc["res"]
    key = ["source", "target", "ligand_complex", "receptor_complex"]
    merged = lr_in[key].merge(res[key + ["lr_probs", "cellchat_pvals"]], on=key, how="left")
    assert len(merged) == len(lr_in) and not merged["cellchat_pvals"].isna().any()
    full = adata.X.tocsr()
    full.sort_indices()
    np.savez_compressed(
        os.path.join(OUT, f"{name}_cellchat.npz"),
        in_indptr=full.indptr.astype(np.int64), in_indices=full.indices.astype(np.int32),
        in_data=full.data.astype(np.float32), in_shape=np.array(full.shape, dtype=np.int64),
        in_var_names=np.array(adata.var_names, dtype=str), in_labels=np.array([str(v) for v in adata.obs[groupby]], dtype=str),
        indptr=X.indptr.astype(np.int64), indices=X.indices.astype(np.int32), data=X.data.astype(np.float32),
        shape=np.array(X.shape, dtype=np.int64), labels=cc["labels"], mat_max=np.float32(lr_in["mat_max"].values[0]),
        cube=cc["cube"].astype(np.float64), ligand_idx=lig, receptor_idx=rec, source_idx=src, target_idx=tgt,
        ligand_trimean=lr_in["ligand_trimean"].values.astype(np.float64),
        receptor_trimean=lr_in["receptor_trimean"].values.astype(np.float64),
        lr_probs=merged["lr_probs"].values.astype(np.float64), pvals=merged["cellchat_pvals"].values.astype(np.float64),
        n_perms=np.int64(n_perms), seed=np.int64(seed), expr_prop=np.float64(expr_prop))
    res.to_csv(os.path.join(OUT, f"{name}_cellchat.csv.gz"), index=False)
    print(f"{name} cellchat: N={X.shape[0]} G={X.shape[1]} P={n_perms} T={len(lig)} nonzero cube {np.mean(cc['cube'] != 0):.3f} "
          f"nonzero lr_probs {np.mean(res['lr_probs'] > 0):.3f} mean pval={res['cellchat_pvals'].mean():.6f}")


def dense_case():
    """700 cells x 300 genes, 6 unequal clusters, high density (many non-zero quartiles), a few negative values
    and explicit zeros: exercises every branch of the sparse order-statistic logic."""
    ad = make_synthetic_adata(700, 300, 6, density=0.45, seed_data=21, seed_labels=22)
    X = ad.X.tocsr().copy()
    X.data[::11] = 0.0
    X.data[5::29] *= -1.0
    return AnnDataLite(X=X, obs=ad.obs, var=ad.var)


def ragged_case():
    """Tiny input with empty cells, explicit stored zeros, unequal clusters and a sub-min_cells cluster."""
    rng = np.random.default_rng(7)
    from liana_py_b200.testing._synth import consensus_genes
    genes = consensus_genes()
    # pick genes that take part in complexes so min-subunit logic is exercised
    res = pd.read_csv(os.path.join(ROOT, "liana_py_b200", "resource", "consensus.csv"))
    cplx = res[res["ligand"].str.contains("_") | res["receptor"].str.contains("_")].head(40)
    chosen = set()
    for col in ("ligand", "receptor"):
        for e in cplx[col]:
            chosen.update(e.split("_"))
    simple = res[~(res["ligand"].str.contains("_") | res["receptor"].str.contains("_"))].head(30)
    chosen.update(simple["ligand"]); chosen.update(simple["receptor"])
    chosen = np.array(sorted(chosen))
    n, g = 90, len(chosen)
    dense = (rng.random((n, g)) < 0.25) * np.log1p(rng.gamma(1.5, 2.0, size=(n, g)))
    dense[5] = 0; dense[40] = 0                     # empty cells (kept by the reference, _pre.py:130-133)
    X = sparse.csr_matrix(dense.astype(np.float32))
    # add explicit stored zeros: they count towards props (getnnz) but not means
    X = X.tolil(); X = X.tocsr()
    X.data[::17] = 0.0
    labels = np.array(["a"] * 40 + ["b"] * 25 + ["c"] * 22 + ["tiny"] * 3)   # 'tiny' < min_cells=5 -> dropped
    perm = rng.permutation(n)
    X = X[perm]; labels = labels[perm]
    order = rng.permutation(g)
    obs = pd.DataFrame({"cluster": pd.Categorical(labels)}, index=[f"c{i}" for i in range(n)])
    var = pd.DataFrame(index=pd.Index(chosen[order]))
    return AnnDataLite(X=X[:, order].tocsr(), obs=obs, var=var)


def save_rank_aggregate(name, adata, groupby, n_perms, seed, expr_prop, ref, **extra):
    """li.mt.rank_aggregate of the reference (default consensus: cellphonedb, connectome, logfc, natmi,
    singlecellsignalr; `liana/method/__init__.py:13-14`) -> final frame as CSV.  `sc.pp.scale` is the only
    scanpy call on this path and is restated in oracle/ref_shim.py."""
    tag = "".join(f"_{k}-{v}" for k, v in sorted(extra.items())) + ("_noperms" if n_perms is None else "")
    res = ref.rank_aggregate(adata.copy(), groupby=groupby, use_raw=False, n_perms=n_perms, seed=seed,
                             expr_prop=expr_prop, inplace=False, verbose=False, **extra)
    res.to_csv(os.path.join(OUT, f"{name}_rank_aggregate{tag}.csv.gz"), index=False)
    print(f"{name} rank_aggregate{tag}: {res.shape}, columns {list(res.columns)[4:]}")


def save_reference_csv_fixture():
    """The reference's own golden frame for rank_aggregate (liana/tests/data/aggregate_rank_rest.csv, produced
    by its authors on pbmc68k_reduced with n_perms=2; asserted by liana/tests/test_rank_aggregate.py:41-47).
    The dataset is not available offline, but the frame pins the consensus step: recomputing both rank columns
    from its score columns must reproduce them."""
    from oracle.ref_shim import REFERENCE_ROOT
    src = os.path.join(REFERENCE_ROOT, "liana", "tests", "data", "aggregate_rank_rest.csv")
    df = pd.read_csv(src, index_col=0)
    df.to_csv(os.path.join(OUT, "ref_aggregate_rank_rest.csv.gz"), index=False)
    print("ref_aggregate_rank_rest:", df.shape)


def save_variants(ref):
    adata, variants = variant_inputs()
    for name, kw in variants.items():
        res = ref.cellphonedb(adata.copy(), groupby="cluster", n_perms=16, seed=7, expr_prop=0.1, inplace=False,
                              verbose=False, **kw)
        res.to_csv(os.path.join(OUT, f"variant_{name}_cellphonedb.csv.gz"), index=False)
        print(f"variant {name}: {res.shape}, mean pval {res['cellphone_pvals'].mean():.4f}")
    # MuData input: two modalities joined by the reference's own `mdata_to_anndata`
    mdata, mkw = mdata_inputs()
    res = ref.cellphonedb(mdata, groupby="cluster", n_perms=16, seed=7, expr_prop=0.1, inplace=False, verbose=False,
                          use_raw=True, mdata_kwargs=mkw)          # use_raw is overridden for MuData (`_liana_pipe.py:128`)
    res.to_csv(os.path.join(OUT, "variant_mdata_cellphonedb.csv.gz"), index=False)
    print(f"variant mdata: {res.shape}, mean pval {res['cellphone_pvals'].mean():.4f}")
    # cells without a cluster label
    adata = nanlabel_inputs()
    res = ref.cellphonedb(adata.copy(), groupby="cluster", n_perms=16, seed=7, expr_prop=0.1, inplace=False, verbose=False,
                          use_raw=False)
    res.to_csv(os.path.join(OUT, "variant_nanlabels_cellphonedb.csv.gz"), index=False)
    print(f"variant nanlabels: {res.shape}, mean pval {res['cellphone_pvals'].mean():.4f}")
    res = ref.rank_aggregate(adata.copy(), groupby="cluster", n_perms=16, seed=7, expr_prop=0.1, inplace=False, verbose=False,
                             use_raw=False)
    res.to_csv(os.path.join(OUT, "variant_nanlabels_rank_aggregate.csv.gz"), index=False)
    print(f"variant nanlabels rank_aggregate: {res.shape}")
    # return_all_lrs (needs the pinned-pandas emulation of oracle/ref_shim.py::_install_chained_fillna_compat)
    adata, _ = variant_inputs()
    for name, fn in (("cellphonedb", ref.cellphonedb), ("natmi", ref.natmi_mod.natmi), ("rank_aggregate", ref.rank_aggregate)):
        res = fn(adata.copy(), groupby="cluster", n_perms=16, seed=7, expr_prop=0.1, inplace=False, verbose=False,
                 use_raw=False, return_all_lrs=True)
        res.to_csv(os.path.join(OUT, f"variant_allrs_{name}.csv.gz"), index=False)
        print(f"variant allrs {name}: {res.shape}, kept {int(res['lrs_to_keep'].sum()) if 'lrs_to_keep' in res else '-'}")
    # by_sample of the reference (`sc/_Method.py:92-159`)
    for name, fn in (("cellphonedb", ref.cellphonedb), ("rank_aggregate", ref.rank_aggregate)):
        res = fn.by_sample(sample_inputs(), sample_key="sample", groupby="cluster", n_perms=16, seed=7, expr_prop=0.1,
                           inplace=False, verbose=False, use_raw=False)
        res.to_csv(os.path.join(OUT, f"variant_bysample_{name}.csv.gz"), index=False)
        print(f"variant bysample {name}: {res.shape}")
    # consensus with a non-default log base and no expression filter (every row scored); magnitude-only consensus
    adata, _ = variant_inputs()
    for tag, kw in (("base2_all", dict(base=2.0, expr_prop=0.0)), ("magnitude_only", dict(consensus_opts=["Magnitude"], expr_prop=0.1))):
        res = ref.rank_aggregate(adata.copy(), groupby="cluster", n_perms=16, seed=3, inplace=False, verbose=False,
                                 use_raw=False, **kw)
        res.to_csv(os.path.join(OUT, f"variant_{tag}_rank_aggregate.csv.gz"), index=False)
        print(f"variant {tag} rank_aggregate: {res.shape}")
    # a custom consensus that contains CellChat: its null scales the shared matrix in place (`_get_mean_perms.py:43-44`),
    # so geometric_mean - which runs after it - is shuffled on X / mat_max (the drop-in's `reference_quirks=True`)
    adata, _ = variant_inputs()
    RA = ref.rank_aggregate_mod
    agg = RA.AggregateClass(RA._rank_aggregate_meta, methods=[ref.cellphonedb, ref.cellchat, ref.geometric_mean])
    res = agg(adata.copy(), groupby="cluster", n_perms=16, seed=3, inplace=False, verbose=False, use_raw=False)
    res.to_csv(os.path.join(OUT, "variant_cellchat_consensus_rank_aggregate.csv.gz"), index=False)
    print(f"variant cellchat consensus: {res.shape}, mean gmean pval {res['gmean_pvals'].mean():.4f}")
    # tied subunits: the reference's tie rule decides which subunit represents a complex
    adata, n_tied = ties_inputs()
    for name, fn in (("cellphonedb", ref.cellphonedb), ("cellchat", ref.cellchat)):
        res = fn(adata.copy(), groupby="cluster", n_perms=16, seed=7, expr_prop=0.1, inplace=False, verbose=False, use_raw=False)
        res.to_csv(os.path.join(OUT, f"variant_ties_{name}.csv.gz"), index=False)
        print(f"variant ties {name}: {res.shape} ({n_tied} complexes with copied subunits)")
    # negative values, integer cluster names
    adata = signed_inputs()
    for name in ("cellphonedb", "cellchat", "geometric_mean"):
        res = getattr(ref, name)(adata.copy(), groupby="cluster", n_perms=16, seed=7, expr_prop=0.1, inplace=False,
                                 verbose=False, use_raw=False)
        res.to_csv(os.path.join(OUT, f"variant_signed_{name}.csv.gz"), index=False)
        print(f"variant signed {name}: {res.shape}")


def save_cellchat(ref):
    save_cellchat_case("small300", make_synthetic_adata(300, 400, 5, density=0.15, seed_data=11, seed_labels=12),
                       "cluster", 32, 42, 0.1, ref)
    save_cellchat_case("dense700", dense_case(), "cluster", 24, 9, 0.1, ref)
    save_cellchat_case("ragged90", ragged_case(), "cluster", 16, 5, 0.05, ref)


def save_scseqcomm(ref):
    """li.mt.scseqcomm of the reference (non-permutation: cluster mean / std of the whole prepared matrix, normal cdf
    of the gene means, `liana/method/sc/_scseqcomm.py`, `_liana_pipe.py:466-499`) on three inputs -> final frames."""
    cases = {"small300": (make_synthetic_adata(300, 400, 5, density=0.15, seed_data=11, seed_labels=12), 0.1),
             "dense700": (dense_case(), 0.1), "ragged90": (ragged_case(), 0.05)}
    for name, (adata, ep) in cases.items():
        res = ref.scseqcomm(adata.copy(), groupby="cluster", use_raw=False, expr_prop=ep, inplace=False, verbose=False)
        res.to_csv(os.path.join(OUT, f"{name}_scseqcomm.csv.gz"), index=False)
        print(f"{name} scseqcomm: {res.shape}, mean inter_score {res['inter_score'].mean():.6f}")


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = load_reference()
    if "--variants-only" in sys.argv:
        save_variants(ref)
        return
    if "--cellchat-only" in sys.argv:
        save_cellchat(ref)
        return
    if "--scseqcomm-only" in sys.argv:
        save_scseqcomm(ref)
        return
    save_variants(ref)
    save_cellchat(ref)
    save_scseqcomm(ref)
    save_reference_csv_fixture()
    save_rank_aggregate("toy700", make_synthetic_adata(700, 765, 10, density=0.10), "cluster", 100, 1337, 0.1, ref)
    small = make_synthetic_adata(300, 400, 5, density=0.15, seed_data=11, seed_labels=12)
    save_rank_aggregate("small300", small, "cluster", 32, 42, 0.1, ref)
    save_rank_aggregate("small300", small, "cluster", 32, 42, 0.1, ref, aggregate_method="mean")
    save_rank_aggregate("ragged90", ragged_case(), "cluster", 16, 5, 0.05, ref)
    # LAST: with n_perms=None the reference sets `cellphonedb.specificity = None` on the shared method object
    # (`_liana_pipe.py:420-422`) and every later permutation call in the same process would fail
    save_rank_aggregate("small300", small, "cluster", None, 42, 0.1, ref)
    if "--rank-only" in sys.argv:
        return
    # config 1 stand-in (BASELINE.json configs[0]): 700 cells x 765 genes, 10 clusters, P=100
    save_case("toy700", make_synthetic_adata(700, 765, 10, density=0.10), "cluster", 100, 1337, 0.1, ref,
              keep_cube_perms=8)
    save_case("small300", make_synthetic_adata(300, 400, 5, density=0.15, seed_data=11, seed_labels=12),
              "cluster", 32, 42, 0.1, ref)
    save_case("ragged90", ragged_case(), "cluster", 16, 5, 0.05, ref)


if __name__ == "__main__":
    main()
