"""CPU: the JSON contract of `bench.py --impl reference` (the arm the driver times beside the engine: the reference's CPU
algorithm on the box's host cores, no GPU needed) on the smoke-size workload: ONE JSON line on stdout with the keys the
driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny",
                          "--steps", "1", "--warmup", "0", "--cpu-budget", "5"], capture_output=True, text=True, timeout=600,
                         cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines                                      # only the JSON line reaches stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "perm*triples/s"
    for k in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config"):
        assert k in d, k
    assert d["value"] > 0 and d["config"]["workload"].startswith("tiny")
    cb = d["cpu_baseline"]
    # the reference's own modules are timed wherever they can be found (/root/reference here, oracle/_ref on the GPU box)
    from oracle import ref_shim
    assert cb["kind"] == ("reference" if ref_shim.reference_available() else "port")
    assert cb["cores"] >= 1 and "sample" in cb and cb["value"] == d["value"]
    assert cb["extrapolated"] is True and cb["measured_perms"] >= 1 and cb["measured_seconds"] > 0
    for k in ("cells", "genes", "genes_in_resource", "clusters", "nnz", "nnz_full", "n_perms", "triples_kept"):
        assert k in d["config"], k
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_a_failing_secondary_section_does_not_cost_the_line():
    """`bench._guarded`: the sections beside the primary numbers (configs[3] / [4], stress, CellChat, ...) record an
    exception under their key instead of aborting the run before the JSON line is printed."""
    import bench
    assert bench._guarded("ok", lambda: {"value": 1}) == {"value": 1}

    def boom():
        raise RuntimeError("x" * 1000)

    out = bench._guarded("stress_expr_prop0", boom)
    assert out["section"] == "stress_expr_prop0" and out["error"].startswith("RuntimeError: xxx") and len(out["error"]) <= 300
    json.dumps(out)
