"""Ligand-receptor resources.

The 17 tables of the reference's ``liana/resource/omni_resource.csv`` (consensus, mouseconsensus, cellphonedb,
cellchatdb, ...) are bundled as gene-symbol CSVs (data, derived by ``scripts/extract_resource.py`` the way
``liana/resource/select_resource.py:9-40`` selects them).  Any other resource is passed as a DataFrame via
``resource=`` or as ``interactions=[(ligand, receptor), ...]`` - same contract as
``liana/resource/select_resource.py:57-82``.
"""
from __future__ import annotations

import os

import pandas as pd

_DIR = os.path.dirname(os.path.abspath(__file__))
_PARSED = {}       # resource name -> parsed frame (callers get copies; `by_sample` asks once per sample)


def show_resources():
    return sorted(f[:-4] for f in os.listdir(_DIR) if f.endswith(".csv"))


def select_resource(resource_name: str = "consensus") -> pd.DataFrame:
    """DataFrame with ['ligand', 'receptor'] columns (complexes as 'A_B')."""
    name = str(resource_name).lower()
    path = os.path.join(_DIR, f"{name}.csv")
    if not os.path.exists(path):
        raise ValueError(f"Resource {name} not found. Please choose from {show_resources()} "
                         "or pass a DataFrame via `resource=`.")
    if name not in _PARSED:
        _PARSED[name] = pd.read_csv(path, index_col=False)
        # identity of an UNMODIFIED bundled table (attrs survive .copy()): lets the table caches skip hashing the content
        _PARSED[name].attrs["lrperm_resource_key"] = ("bundled", name, len(_PARSED[name]))
    return _PARSED[name].copy()


def handle_resource(interactions=None, resource=None, resource_name=None, verbose=False) -> pd.DataFrame:
    """Precedence interactions > resource > resource_name, validation as the reference."""
    if interactions is not None:
        if not isinstance(interactions, list) or any(len(item) != 2 for item in interactions):
            raise ValueError("'interactions' should be a list of tuples in the format [(x1, y1), (x2, y2), ...].")
        return pd.DataFrame(list(dict.fromkeys(map(tuple, interactions))), columns=["ligand", "receptor"])
    if resource is not None:
        if not isinstance(resource, pd.DataFrame) or "ligand" not in resource.columns or "receptor" not in resource.columns:
            raise ValueError("If 'interactions' is None, 'resource' must be a valid DataFrame "
                             "with columns 'ligand' and 'receptor'.")
        out = resource[["ligand", "receptor"]].dropna(subset=["ligand", "receptor"]).drop_duplicates()
        return out.reset_index(drop=True)
    if resource_name is None:
        raise ValueError("If 'interactions' and 'resource' are both None, 'resource_name' must be provided.")
    return select_resource(resource_name)
