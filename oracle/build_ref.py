"""TEST INFRASTRUCTURE ONLY -- recipe that builds ``oracle/_ref``: the reference's own hot-path modules,
byte-compiled from the sources where they lie under ``/root/reference``.

    python oracle/build_ref.py            # no-op (exit 0) when /root/reference is absent

The reference is pure Python, so "building" it means ``py_compile``: every ``.py`` of the three packages the
path imports (``utils``, ``models``, ``data_loaders``) becomes a source-less byte-code file (``.rbc``: the ``.pyc``
format under a neutral extension, so that no snapshot rule for Python caches drops it) under ``oracle/_ref/`` with
the same package layout; ``oracle/ref_loader.py`` imports them through a small meta-path finder.  No reference source is copied into the repository: ``oracle/_ref/`` holds compiled
artefacts only, is git-ignored (like the built ``.so``) and NOT gpurun-ignored, so it travels to the GPU box,
where ``/root/reference`` does not exist.  ``oracle/ref_loader.py`` imports from ``/root/reference`` when it is
there and from ``oracle/_ref`` otherwise; ``bench.py``'s CPU legs then time the reference itself
(``cpu_baseline.kind == "reference"``) instead of the NumPy port.

``__graft_entry__.build()`` runs this (building the checker is not using it)."""
from __future__ import annotations

import os
import py_compile
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("SSDGEOM_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")
PACKAGES = ("utils", "models", "data_loaders")


def build(verbose: bool = False) -> str | None:
    if not os.path.isfile(os.path.join(SRC, "utils", "bbox.py")):
        return OUT if os.path.isfile(os.path.join(OUT, "utils", "bbox.rbc")) else None
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    n = 0
    for pkg in PACKAGES:
        for dirpath, _dirs, files in os.walk(os.path.join(SRC, pkg)):
            rel = os.path.relpath(dirpath, SRC)
            for name in files:
                if not name.endswith(".py"):
                    continue
                dst_dir = os.path.join(OUT, rel)
                os.makedirs(dst_dir, exist_ok=True)
                # dfile: the path tracebacks will name -- the real location of the source
                py_compile.compile(os.path.join(dirpath, name), cfile=os.path.join(dst_dir, name[:-3] + ".rbc"),
                                   dfile=os.path.join(SRC, rel, name), doraise=True)
                n += 1
    with open(os.path.join(OUT, "BUILT_FROM"), "w") as f:
        f.write("%s (python %d.%d, %d modules byte-compiled by oracle/build_ref.py)\n" %
                (SRC, sys.version_info[0], sys.version_info[1], n))
    if verbose:
        print("oracle/_ref: %d modules byte-compiled from %s" % (n, SRC))
    return OUT


if __name__ == "__main__":
    print(build(verbose=True))
