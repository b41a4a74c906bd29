"""TEST INFRASTRUCTURE ONLY -- import the *unmodified* reference source from
``/root/reference`` under the NumPy ``tensorflow`` shim (oracle/tf_shim.py).

``/root/reference`` exists only in the build container.  On the GPU box the same modules are imported from
``oracle/_ref`` -- source-less byte-code files byte-compiled (``.rbc``) from ``/root/reference`` by ``oracle/build_ref.py``
(git-ignored build artefacts that travel with the snapshot, like the built ``.so``).  It is used by

* ``oracle/make_golden.py`` to generate ``tests/golden/*.npz``,
* ``tests/test_oracle_vs_reference.py`` (skipped when the reference is absent)
  to pin ``oracle/ssd_oracle.py`` against the real thing, and
* ``bench.py``'s CPU legs (``--impl reference`` and ``cpu_baseline``), which time the reference's own
  ``match_bbox`` / ``apply_anchor_box`` / ``_ssd_loss`` when either location is present
  (``cpu_baseline.kind == "reference"``), else the NumPy port (``"port"``).

Reference entry points exposed (file:line in /root/reference):
  utils/bbox.py:6    iou
  utils/bbox.py:28   iou_n
  utils/bbox.py:44   match_bbox
  utils/bbox.py:94   apply_anchor_box
  models/ssd_model.py:173  SSDObjectDetectionModel._build_prior_box
  models/ssd_model.py:341  SSDObjectDetectionModel._ssd_loss
  data_loaders/ssd/make_dataset.py:37  SSDDataLoader._coco2ssd (resize + relative boxes)
"""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.machinery
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SSDGEOM_REFERENCE_ROOT", "/root/reference")
COMPILED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def root() -> str | None:
    """Where the reference's modules can be imported from: its source tree, else the byte-compiled copy."""
    if os.path.isfile(os.path.join(REFERENCE_ROOT, "utils", "bbox.py")):
        return REFERENCE_ROOT
    if os.path.isfile(os.path.join(COMPILED_ROOT, "utils", "bbox.rbc")):
        return COMPILED_ROOT
    return None


class _CompiledFinder(importlib.abc.MetaPathFinder):
    """Imports ``utils`` / ``models`` / ``data_loaders`` from the byte-code files oracle/build_ref.py wrote."""

    def __init__(self, base):
        self.base = base

    def find_spec(self, fullname, path=None, target=None):
        parts = fullname.split(".")
        if parts[0] not in ("utils", "models", "data_loaders"):
            return None
        where = os.path.join(self.base, *parts)
        init = os.path.join(where, "__init__.rbc")
        if os.path.isfile(init):
            return importlib.util.spec_from_file_location(
                fullname, init, loader=importlib.machinery.SourcelessFileLoader(fullname, init),
                submodule_search_locations=[where])
        if os.path.isfile(where + ".rbc"):
            return importlib.util.spec_from_file_location(
                fullname, where + ".rbc", loader=importlib.machinery.SourcelessFileLoader(fullname, where + ".rbc"))
        return None


def available() -> bool:
    return root() is not None


_cache = {}


def load():
    """Return a namespace with the reference's hot-path callables."""
    if "ns" in _cache:
        return _cache["ns"]
    where = root()
    if where is None:
        raise RuntimeError("reference not present at %s nor byte-compiled under %s" % (REFERENCE_ROOT, COMPILED_ROOT))
    from . import tf_shim

    tf_shim.install()
    # The reference imports its packages as top-level names (utils, models, data_loaders).
    # Import them under those names from REFERENCE_ROOT, then restore sys.path.
    for name in list(sys.modules):
        if name.split(".")[0] in ("utils", "models", "data_loaders"):
            raise RuntimeError("a module named %r is already imported; cannot load the reference" % name)
    hook = _CompiledFinder(where) if where == COMPILED_ROOT else None
    if hook is not None:
        sys.meta_path.insert(0, hook)
    else:
        sys.path.insert(0, where)
    try:
        bbox = importlib.import_module("utils.bbox")
        ssd_model = importlib.import_module("models.ssd_model")
    finally:
        if hook is not None:
            sys.meta_path.remove(hook)
        else:
            sys.path.remove(where)
    model_cls = ssd_model.SSDObjectDetectionModel

    def build_prior_box(size_list, input_size=300):
        fake_self = types.SimpleNamespace(cfg=types.SimpleNamespace(input_shape=(input_size, input_size, 3)))
        return model_cls._build_prior_box(fake_self, size_list)

    ssd_loader = sys.modules["data_loaders.ssd.make_dataset"].SSDDataLoader

    def coco2ssd(image, cls, box, train_resize=(300, 300)):
        """(image, cls, box) -> (resized image, cls, box / [w,h,w,h]); mutates ``box`` in place like the reference."""
        return ssd_loader._coco2ssd(types.SimpleNamespace(_train_resize=train_resize), (image, cls, box))

    ns = types.SimpleNamespace(
        coco2ssd=coco2ssd,
        iou=bbox.iou,
        iou_n=bbox.iou_n,
        match_bbox=bbox.match_bbox,
        apply_anchor_box=bbox.apply_anchor_box,
        build_prior_box=build_prior_box,
        ssd_loss=model_cls._ssd_loss,
        bbox_module=bbox,
        model_module=ssd_model,
        root=where,
    )
    _cache["ns"] = ns
    return ns
