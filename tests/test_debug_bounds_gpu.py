"""GPU: the kernels' own bounds checks (compute-sanitizer is closed on the B200 pool, profiles/r02_sanitizer_closed.md).
`liblrperm_debug.so` is the same library compiled with -DLRPERM_BOUNDS_CHECK: every index the hot kernels compute -
shared-memory bin offsets, accumulator columns, cells, chunk ranges, partial-sum slots, cube rows, trimean planes - is
tested on the device against the engine's extents and reported through `lrperm_debug_bounds`.  A worker process loads
that build (LRPERM_LIB) and runs the golden fixtures (EXACT, FAST, trimean) and ragged / tiled / filtered inputs; no
check may fire.  The production library contains none of the checks and says so."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_production_library_has_no_bounds_checks(engine):
    from liana_py_b200 import _engine as E
    with pytest.raises(E.EngineError, match="LRPERM_BOUNDS_CHECK"):
        engine.debug_bounds()


def test_kernels_stay_in_bounds_under_the_debug_build():
    from liana_py_b200 import build
    # prebuilt in-tree by __graft_entry__.build() (it travels with the snapshot); compiled here only when it is missing
    lib = build.OUT_DEBUG if os.path.exists(build.OUT_DEBUG) else build.build(debug=True)
    env = dict(os.environ, LRPERM_LIB=lib)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "_debug_bounds_worker.py")], env=env, capture_output=True,
                         text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "BOUNDS_OK" in out.stdout, out.stdout[-2000:]
