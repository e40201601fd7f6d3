"""Container-only: derive the `consensus` ligand-receptor table (data, not code) from the
reference's omni_resource.csv (`liana/resource/select_resource.py:9-40` selects it the same
way: filter resource=='consensus', keep source_genesymbol/target_genesymbol)."""
import sys
import pandas as pd

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/liana/resource/omni_resource.csv"
names = sys.argv[2:] or ["consensus"]
df = pd.read_csv(src, index_col=False)
for name in names:
    sub = df[df["resource"] == name][["source_genesymbol", "target_genesymbol"]]
    sub = sub.rename(columns={"source_genesymbol": "ligand", "target_genesymbol": "receptor"})
    out = f"liana_py_b200/resource/{name}.csv"
    sub.to_csv(out, index=False)
    print(name, sub.shape, "->", out)
