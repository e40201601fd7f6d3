"""Where the consensus of an all-rows (expr_prop=0) run departs from the reference's frame: per column statistics."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, pandas as pd
import liana_py_b200 as li
from variants import variant_inputs
KEY = ["source", "target", "ligand_complex", "receptor_complex"]
adata, _ = variant_inputs()
res = li.mt.rank_aggregate(adata, base=2.0, expr_prop=0.0, groupby="cluster", n_perms=16, seed=3, inplace=False,
                           perm_source="reference", use_raw=False)
gold = pd.read_csv("tests/golden/variant_base2_all_rank_aggregate.csv.gz", float_precision="round_trip")
a = res.sort_values(KEY).reset_index(drop=True); b = gold.sort_values(KEY).reset_index(drop=True)
for c in [c for c in gold.columns if c not in KEY]:
    x, y = a[c].values.astype(np.float64), b[c].values.astype(np.float64)
    rel = np.abs(x - y) / np.maximum(np.abs(y), 1e-12)
    print(f"{c:18s} exact {np.mean(x == y):.4f}  rel<=1e-4 {np.mean(rel <= 1e-4):.4f}  max rel {rel.max():.3e}  max abs {np.abs(x - y).max():.3e}"
          f"  distinct values ours/ref {len(np.unique(x))}/{len(np.unique(y))}")
bad = np.flatnonzero(np.abs(a["specificity_rank"].values - b["specificity_rank"].values) / np.maximum(b["specificity_rank"].values, 1e-12) > 1e-4)
print(len(bad), "rows off in specificity_rank; sample:")
print(pd.concat([a.loc[bad[:6], KEY + ["scaled_weight", "lr_logfc", "spec_weight", "specificity_rank"]],
                 b.loc[bad[:6], ["scaled_weight", "lr_logfc", "spec_weight", "specificity_rank"]].add_suffix("_ref")], axis=1).to_string())
