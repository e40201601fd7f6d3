"""TEST / BENCH INFRASTRUCTURE ONLY - recipe that makes the UNMODIFIED reference travel to the GPU box.

    python oracle/make_ref.py            (also run by __graft_entry__.build() when /root/reference is present)

Copies the reference's own Python package tree (`/root/reference/liana`, minus its tests) byte for byte into
`oracle/_ref/liana`.  `oracle/_ref/` is git-ignored (the reference's sources never enter this repository's history)
but not gpurun-ignored, so the copy ships with the snapshot like the built `.so` files.  `oracle/ref_shim.py` loads
the hot-path modules from it (`liana/method/_pipe_utils/_get_mean_perms.py`, `liana/method/sc/_liana_pipe.py`,
`_cellphonedb.py`, ... - the list is in the shim's header) when `/root/reference` itself does not exist, which is
how `bench.py --impl reference` and `cpu_baseline` time the reference's own code (`kind: "reference"`) there.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("LIANA_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")


def make(force: bool = False) -> str | None:
    src = os.path.join(SRC, "liana")
    if not os.path.isdir(src):
        return None                                  # GPU box: the copy made in the build container is used as it is
    dst = os.path.join(DST, "liana")
    ignore = shutil.ignore_patterns("tests", "__pycache__", "*.pyc")
    stamp = os.path.join(DST, "STAMP.json")
    if os.path.isdir(dst) and not force:
        try:
            with open(stamp) as f:
                if json.load(f).get("source_digest") == tree_digest_filtered(src):
                    return dst
        except Exception:
            pass
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    os.makedirs(DST, exist_ok=True)
    shutil.copytree(src, dst, ignore=ignore)
    with open(stamp, "w") as f:
        json.dump({"source": src, "source_digest": tree_digest_filtered(dst),
                   "what": "unmodified copy of the reference's liana/ package (tests excluded); see oracle/make_ref.py"}, f)
    return dst


def tree_digest_filtered(src: str) -> str:
    """Digest of the source tree over the same files the copy holds (tests excluded)."""
    h = hashlib.sha256()
    for base, dirs, files in os.walk(src):
        dirs[:] = sorted(d for d in dirs if d not in ("tests", "__pycache__"))
        for f in sorted(files):
            if f.endswith(".py") or f.endswith(".csv"):
                p = os.path.join(base, f)
                h.update(os.path.relpath(p, src).encode())
                with open(p, "rb") as fh:
                    h.update(fh.read())
    return h.hexdigest()


if __name__ == "__main__":
    out = make(force="--force" in sys.argv)
    print(out if out else f"{SRC}/liana not found: nothing copied (an existing oracle/_ref is kept)")
