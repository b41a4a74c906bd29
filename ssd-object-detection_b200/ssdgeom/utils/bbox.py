"""Drop-in for the hot part of the reference's ``utils/bbox.py`` (lines 6-101): same names,
argument meaning, return dtypes and error behaviour; every function runs on the GPU through
libssdgeom.so.  Host (NumPy) inputs are uploaded and results downloaded; device inputs
(``__cuda_array_interface__``) are used in place.

Not provided: ``draw_bbox`` (utils/bbox.py:104-147, OpenCV drawing -- out of scope)."""
from __future__ import annotations

import numpy as np

from .. import device as D
from .. import ops


def _boxes(x):
    a = np.asarray(x)
    if a.dtype not in (np.float32, np.float64):
        a = a.astype(np.float64)           # NumPy's promotion for Python / integer inputs
    return np.ascontiguousarray(a)


def iou(bbox_1, bbox_2):
    """utils/bbox.py:6-25: IoU of cxcywh boxes (scalars-of-4 or [n,4] arrays), intersection
    extents clamped at 0.0, denominator + 1e-10."""
    b1, b2 = _boxes(bbox_1), _boxes(bbox_2)
    single = b1.ndim == 1
    # the reference indexes bbox[0..3]: a [4] box or a [4,n] stack of columns
    b1 = b1.reshape(4, -1).T if b1.ndim == 2 else b1.reshape(1, 4)
    b2 = b2.reshape(4, -1).T if b2.ndim == 2 else b2.reshape(1, 4)
    out = ops.iou_pairs(np.ascontiguousarray(b1), np.ascontiguousarray(b2), use_eps_clamp=False).to_host()
    return out[0] if single else out


def iou_n(n_bbox_1, n_bbox_2):
    """utils/bbox.py:28-41: paired rows of two [n,4] arrays, extents clamped at 1e-10."""
    b1, b2 = _boxes(n_bbox_1), _boxes(n_bbox_2)
    return ops.iou_pairs(b1, b2, use_eps_clamp=True).to_host()


def match_bbox(cls, bbox, default_box, thresh=0.5, return_match=False):
    """utils/bbox.py:44-91: returns (labeled_cls int32[A], labeled_boxes float32[A,4], mask bool[A])."""
    gt_box, priors = _boxes(bbox), _boxes(default_box)
    gt_cls = np.asarray(cls)
    n_gt, n_pr = gt_box.shape[0], priors.shape[0]
    assert n_gt <= n_pr, "number of default boxes should greater than the number of targets"
    assert thresh > 0.0, "thresh should greater than zero"
    if n_gt == 0:
        raise ValueError("attempt to get argmax of an empty sequence")   # np.argmax at utils/bbox.py:72
    # int(target_cls[t]) at :90 truncates toward zero; small integers are exact in float32
    gt_cls32 = np.trunc(gt_cls.astype(np.float64)).astype(np.float32)
    offsets = np.array([0, n_gt], dtype=np.int32)
    want = ("cls", "box", "mask") + (("match",) if return_match else ())
    out = ops.match_encode(gt_box, gt_cls32, offsets, priors, 1, n_gt, float(thresh), want=want)
    res = (out["cls"].to_host()[0], out["box"].to_host()[0], out["mask"].to_host()[0].astype(bool))
    return res + (out["match"].to_host()[0],) if return_match else res


def apply_anchor_box(origin_bbox, default_box):
    """utils/bbox.py:94-101: offset encoding, no variances; float64 unless both inputs are float32."""
    g, d = _boxes(origin_bbox), _boxes(default_box)
    assert np.shape(g) == np.shape(d)
    odt = np.float32 if (g.dtype == np.float32 and d.dtype == np.float32) else np.float64
    return ops.encode(g, d, out_dtype=odt).to_host()


def decode_bbox(loc, default_box, scale=300.0, variances=None):
    """The inverse the reference applies inline at models/ssd_model.py:466-467 (pixels for scale=300).
    ``variances=(v_xy, v_wh)`` (not in the reference): the offsets are multiplied by them first."""
    t = np.ascontiguousarray(np.asarray(loc, dtype=np.float32))
    if variances is not None:
        t = ops.loc_scale(t, variances[0], variances[1])
    return ops.decode(t, _boxes(default_box), scale=scale).to_host()


def match_encode_batch(gt_boxes, gt_cls, gt_offsets, default_box, thresh=0.5, device_out=False, stream=None,
                       variances=None):
    """The reference's per-image generator body (models/ssd_model.py:211-215) for a whole batch:
    CSR ground truth -> (cls int32[B,A], loc float32[B,A,4], mask bool[B,A]).  ``variances=(v_xy, v_wh)`` (not in
    the reference) divides the encoded offsets."""
    off = np.asarray(gt_offsets, dtype=np.int32)
    counts = np.diff(off)
    priors = default_box if D.is_device(default_box) else _boxes(default_box)
    n_pr = int(priors.shape[0])
    assert counts.size > 0 and int(counts.max()) <= n_pr, "number of default boxes should greater than the number of targets"
    assert thresh > 0.0, "thresh should greater than zero"
    gb = gt_boxes if D.is_device(gt_boxes) else _boxes(gt_boxes)
    gc = gt_cls if D.is_device(gt_cls) else np.trunc(np.asarray(gt_cls, dtype=np.float64)).astype(np.float32)
    out = ops.match_encode(gb, gc, off, priors, int(counts.size), int(counts.max()), float(thresh), stream=stream)
    if variances is not None:
        ops.loc_scale(out["loc"], 1.0 / variances[0], 1.0 / variances[1], out=out["loc"], stream=stream)
    if device_out:
        return out["cls"], out["loc"], out["mask"]
    return out["cls"].to_host(stream), out["loc"].to_host(stream), out["mask"].to_host(stream).astype(bool)
