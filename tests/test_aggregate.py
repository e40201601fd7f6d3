"""CPU: the consensus step (average ranks + RobustRankAggregate / mean rank) against known answers.

* `ref_aggregate_rank_rest.csv.gz` is the reference's OWN golden frame
  (liana/tests/data/aggregate_rank_rest.csv, asserted by liana/tests/test_rank_aggregate.py:41-47):
  recomputing both rank columns from its score columns must reproduce the stored ranks.
* `*_rank_aggregate*.csv.gz` were produced in the build container by the unmodified reference
  (scripts/make_golden.py) on the synthetic fixtures.
"""
import os

import numpy as np
import pandas as pd
import pytest

import liana_py_b200 as li
from liana_py_b200.method._aggregate import aggregate, rank_aggregate_scores, robust_rank_aggregate

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KEY = ["source", "target", "ligand_complex", "receptor_complex"]


def _read(name):
    return pd.read_csv(os.path.join(GOLDEN, name), float_precision="round_trip")


def test_specs_match_reference():
    # liana/tests/test_rank_aggregate.py:24-38
    ra = li.mt.rank_aggregate
    assert ra.specificity_specs == {"CellPhoneDB": ("cellphone_pvals", True), "Connectome": ("scaled_weight", False),
                                    "log2FC": ("lr_logfc", False), "NATMI": ("spec_weight", False)}
    assert ra.magnitude_specs == {"CellPhoneDB": ("lr_means", False), "Connectome": ("expr_prod", False),
                                  "NATMI": ("expr_prod", False), "SingleCellSignalR": ("lrscore", False)}
    assert ra.magnitude == "magnitude_rank" and ra.specificity == "specificity_rank"
    assert ra.method_name == "Rank_Aggregate"


# The reference's own CSV was written under its pinned scipy (float64 ranks) -> reproduced to 1e-12.  The frames
# generated in this container ran the same reference code on scipy 1.18, whose rankdata keeps float32 for
# float32 scores, so their stored ranks carry ~1e-7 of float32 rounding (oracle caveat).
@pytest.mark.parametrize("fname,method,rtol", [
    ("ref_aggregate_rank_rest.csv.gz", "rra", 1e-12), ("toy700_rank_aggregate.csv.gz", "rra", 1e-6),
    ("small300_rank_aggregate.csv.gz", "rra", 1e-6), ("ragged90_rank_aggregate.csv.gz", "rra", 1e-6),
    ("small300_rank_aggregate_aggregate_method-mean.csv.gz", "mean", 1e-6)])
def test_ranks_reproduce_reference_frames(fname, method, rtol):
    gold = _read(fname)
    scores = gold.drop(columns=["specificity_rank", "magnitude_rank"]).copy()
    out = aggregate(scores, consensus=li.mt.rank_aggregate, aggregate_method=method)
    out = out.sort_values(KEY).reset_index(drop=True)
    exp = gold.sort_values(KEY).reset_index(drop=True)
    # the reference's own golden CSV stores its float32 scores with 8 digits -> rank ties are preserved
    np.testing.assert_allclose(out["specificity_rank"].values, exp["specificity_rank"].values, rtol=rtol, atol=0)
    np.testing.assert_allclose(out["magnitude_rank"].values, exp["magnitude_rank"].values, rtol=rtol, atol=0)
    # float32 score columns (what the pipeline produces) give the same ranks as their float64 read-back
    f32 = scores.copy()
    for c in f32.columns:
        if f32[c].dtype == np.float64 and c not in ("cellphone_pvals",):
            f32[c] = f32[c].astype(np.float32)
    out32 = aggregate(f32, consensus=li.mt.rank_aggregate, aggregate_method=method).sort_values(KEY).reset_index(drop=True)
    np.testing.assert_allclose(out32["magnitude_rank"].values, exp["magnitude_rank"].values, rtol=max(rtol, 1e-9), atol=0)
    assert np.all(np.diff(aggregate(scores, li.mt.rank_aggregate, method)["magnitude_rank"].values) >= 0)


def test_noperms_frame_only_has_magnitude():
    gold = _read("small300_rank_aggregate_noperms.csv.gz")
    assert "specificity_rank" not in gold.columns and "cellphone_pvals" not in gold.columns
    scores = gold.drop(columns=["magnitude_rank"]).copy()
    out = aggregate(scores, li.mt.rank_aggregate, "rra", consensus_opts="Magnitude")
    assert "specificity_rank" not in out.columns
    a = out.sort_values(KEY)["magnitude_rank"].values
    b = gold.sort_values(KEY)["magnitude_rank"].values
    np.testing.assert_allclose(a, b, rtol=1e-6)


def test_rra_properties():
    rng = np.random.default_rng(0)
    n, m = 500, 4
    rmat = np.stack([rng.permutation(n) + 1.0 for _ in range(m)], axis=1)
    rho = robust_rank_aggregate(rmat)
    assert rho.shape == (n,) and np.all((rho >= 0) & (rho <= 1))
    # a row ranked first by every method gets the smallest score; column order is irrelevant
    rmat[0] = 1.0
    rho = robust_rank_aggregate(rmat)
    assert rho.argmin() == 0
    assert np.allclose(rho, robust_rank_aggregate(rmat[:, ::-1]))
    # expr_prod is named by two methods and therefore ranked twice (reference quirk, kept on purpose)
    df = pd.DataFrame({"expr_prod": [3.0, 1.0, 2.0]})
    once = rank_aggregate_scores(df, {"A": ("expr_prod", False)}, "mean")
    twice = rank_aggregate_scores(df, {"A": ("expr_prod", False), "B": ("expr_prod", False)}, "mean")
    assert np.array_equal(np.argsort(once), np.argsort(twice)[::-1])


def test_method_registry_matches_reference_records():
    """`li.mt.show_methods()` / `get_method_scores()` (`liana/method/__init__.py:17-32`) and the MethodMeta records of the
    three permutation methods (`sc/_cellphonedb.py:34-50`, `_geometric_mean.py:27-41`, `_cellchat.py:36-53`)."""
    import liana_py_b200 as li
    meta = li.mt.show_methods()
    assert list(meta["Method Name"]) == ["CellPhoneDB", "Connectome", "log2FC", "NATMI", "SingleCellSignalR",
                                         "Rank_Aggregate", "Geometric Mean", "scSeqComm", "CellChat"]
    scores = li.mt.get_method_scores()
    assert scores["cellphone_pvals"] is True and scores["lr_means"] is False and scores["cellchat_pvals"] is True
    assert scores["lr_probs"] is False and scores["magnitude_rank"] is True
    cc = li.mt.cellchat
    assert (cc.complex_cols, cc.add_cols, cc.magnitude, cc.specificity, cc.permute, cc._kh) == (
        ["ligand_trimean", "receptor_trimean"], ["mat_max"], "lr_probs", "cellchat_pvals", True, 0.5)
    cp = li.mt.cellphonedb
    assert (cp.complex_cols, cp.add_cols, cp.magnitude, cp.specificity, cp.magnitude_ascending, cp.specificity_ascending) == (
        ["ligand_means", "receptor_means"], [], "lr_means", "cellphone_pvals", False, True)


def test_bundled_resources_match_reference_tables():
    """Every resource of the reference's `omni_resource.csv` is bundled as a gene-symbol table (data derived by
    scripts/extract_resource.py) and selected the way `liana/resource/select_resource.py:9-40` selects it
    (case-insensitive name, ValueError for unknown names).  The row-by-row comparison with the reference's file runs
    only where /root/reference exists (the build container)."""
    import os
    import pandas as pd
    import pytest
    import liana_py_b200 as li
    names = li.rs.show_resources()
    assert "consensus" in names and "cellphonedb" in names and "mouseconsensus" in names and len(names) == 17
    assert li.rs.select_resource("Consensus").shape == (4624, 2)
    with pytest.raises(ValueError, match="not found"):
        li.rs.select_resource("nope")
    src = "/root/reference/liana/resource/omni_resource.csv"
    if os.path.exists(src):
        omni = pd.read_csv(src, index_col=False)
        assert sorted(omni["resource"].unique()) == names
        for name in names:
            ref = omni[omni["resource"] == name][["source_genesymbol", "target_genesymbol"]]
            got = li.rs.select_resource(name)
            assert list(got.columns) == ["ligand", "receptor"]
            assert (got["ligand"].values == ref["source_genesymbol"].values).all()
            assert (got["receptor"].values == ref["target_genesymbol"].values).all()
