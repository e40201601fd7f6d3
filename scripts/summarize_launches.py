"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> markdown table per kernel name.
   python scripts/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches.md"""
import collections, csv, re, sys

ENGINE = ("fast2_means", "fast3_means", "fast_means", "exact_means", "slab_", "fast_finalize", "score_count", "stage_stats", "stage_max",
          "mapped_", "pack_csr", "csc_keys", "stored_counts", "copy_bytes", "copy_words", "mark_genes", "resolve_", "unpack_pairs",
          "label_", "narrow_indptr", "widen_indptr", "validate_columns", "build_rows", "cube_export", "cluster_moments", "moments_reduce",
          "rank_", "rra_", "tm_", "vsc_", "scale_labels", "dense_", "row_keep", "new_row", "philox_perms", "staged_cluster")


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"\(.*$", "", name)
    depth, out = 0, []
    for ch in name:                       # drop template arguments
        if ch == "<":
            depth += 1
        elif ch == ">":
            depth -= 1
        elif depth == 0:
            out.append(ch)
    return "".join(out).split("::")[-1].strip() or name[:40]


path = sys.argv[1]
lines = [ln for ln in open(path) if ln.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "s": 1e3, "second": 1e3}
agg = collections.OrderedDict()
for r in rd:
    a = agg.setdefault(short(r[ki]), [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
total = sum(a[1] for a in agg.values())
mine = {k: v for k, v in agg.items() if k.startswith(ENGINE)}
t_mine = sum(v[1] for v in mine.values())
print(f"Launches: {sum(a[0] for a in agg.values())}, summed device time {total:.1f} ms "
      f"(engine kernels {t_mine:.1f} ms; the rest is CUB inside the engine and the torch kernels of the synthetic-data generator).  "
      "Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n")
print("| kernel | launches | device time (ms) | share of all | share of engine kernels |")
print("|---|---:|---:|---:|---:|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if t / total < 0.0005:
        continue
    eng = f"{100 * t / t_mine:.1f} %" if k in mine else "-"
    print(f"| `{k}` | {n} | {t:.2f} | {100 * t / total:.1f} % | {eng} |")
step = {k: agg[k][1] for k in ("fast2_means_kernel", "slab_philox_kernel", "fast_finalize_kernel", "score_count_kernel") if k in agg}
if step:
    s = sum(step.values())
    print("\nThe four kernels of a FAST step: " + ", ".join(f"`{k}` {100 * v / s:.1f} %" for k, v in step.items()) + ".")
