"""TEST INFRASTRUCTURE ONLY -- a NumPy-backed stand-in for the handful of
``tensorflow`` symbols the reference's hot path touches.

TensorFlow is not installed in this image (reference requirements.txt:1 asks for
``tensorflow>=2.4.0``, un-pinned).  The reference's box arithmetic is NumPy
(utils/bbox.py:28-101) and its loss is a dozen eager TF ops
(models/ssd_model.py:341-396).  This shim restates the *published* semantics of
exactly those ops so the reference source can be executed verbatim from
``/root/reference`` in this container to (a) pin the oracle restatement in
``oracle/ssd_oracle.py`` and (b) generate the fixtures under ``tests/golden``.

Nothing under ``ssdgeom`` (the product) imports this file.

Numerics of the restated TF ops (parity for these is "unpinned" by any reference
test):

* ``sparse_softmax_cross_entropy_with_logits``: ``log(sum(exp(x-max))) - (x[label]-max)``
  evaluated in float64 from the float32 logits, rounded once to float32 (the
  dtype TF would hold).
* ``math.top_k``: values sorted descending (stable), float32.
* ``reduce_sum``: accumulated in float64, so the oracle's loss scalars are the
  "true" value the 1e-5 tolerance is measured from.
"""
from __future__ import annotations

import sys
import types
from unittest import mock

import numpy as np


class _T(np.ndarray):
    """ndarray with the two Tensor conveniences the reference calls."""

    def numpy(self):
        return np.asarray(self)


def _t(x):
    return np.asarray(x).view(_T)


def _cast(x, dtype):
    return _t(np.asarray(x).astype(dtype))


def _reduce_sum(x, axis=None):
    a = np.asarray(x)
    if a.dtype.kind == "f":
        return _t(np.sum(a, axis=axis, dtype=np.float64))
    return _t(np.sum(a, axis=axis))


def _sparse_ce(labels, logits):
    x = np.asarray(logits).astype(np.float64)
    lab = np.asarray(labels).astype(np.int64)
    m = x.max(axis=-1, keepdims=True)
    lse = np.log(np.exp(x - m).sum(axis=-1))
    picked = np.take_along_axis(x - m, lab[..., None], axis=-1)[..., 0]
    return _t((lse - picked).astype(np.float32))


def _top_k(x, k):
    a = np.asarray(x)
    k = int(k)
    if k > a.shape[-1]:
        raise ValueError("input must have at least k columns")
    order = np.argsort(-a, kind="stable")[:k]
    return _t(a[order]), _t(order.astype(np.int32))


def _softmax(x, axis=-1):
    a = np.asarray(x).astype(np.float64)
    e = np.exp(a - a.max(axis=axis, keepdims=True))
    return _t((e / e.sum(axis=axis, keepdims=True)).astype(np.float32))


def build_module() -> types.ModuleType:
    tf = types.ModuleType("tensorflow")
    tf.__dict__.update(
        Tensor=_T,
        float32=np.float32, float64=np.float64, int32=np.int32, int64=np.int64, bool=np.bool_,
        Variable=lambda v, dtype=None: _t(np.array(v, dtype=dtype)),
        constant=lambda v, dtype=None: _t(np.array(v, dtype=dtype)),
        maximum=np.maximum, minimum=np.minimum,
        shape=lambda x: np.shape(x),
        boolean_mask=lambda x, m: _t(np.asarray(x)[np.asarray(m).astype(bool)]),
        equal=lambda a, b: _t(np.equal(a, b)),
        zeros_like=lambda x, dtype=None: _t(np.zeros_like(x, dtype=dtype)),
        ones_like=lambda x, dtype=None: _t(np.ones_like(x, dtype=dtype)),
        cast=_cast,
        reduce_sum=_reduce_sum,
        reduce_min=lambda x, axis=None: _t(np.min(x, axis=axis)),
        reduce_max=lambda x, axis=None: _t(np.max(x, axis=axis)),
        reshape=lambda x, s: _t(np.reshape(x, s)),
        abs=lambda x: _t(np.abs(x)),
        logical_and=lambda a, b: _t(np.logical_and(a, b)),
        logical_not=lambda a: _t(np.logical_not(a)),
        argmax=lambda x, axis=None: _t(np.argmax(x, axis=axis)),
        zeros=lambda s, dtype=np.float32: _t(np.zeros(s, dtype=dtype)),
        function=lambda f: f,
    )
    nn = types.ModuleType("tensorflow.nn")
    nn.sparse_softmax_cross_entropy_with_logits = _sparse_ce
    nn.softmax = _softmax
    math_ = types.ModuleType("tensorflow.math")
    math_.top_k = _top_k
    tf.nn, tf.math = nn, math_
    # Everything the hot path does not touch (keras layers/optimizers evaluated at
    # class-definition time models/ssd_model.py:26-27, tf.summary, tf.data) is a mock.
    tf.keras = mock.MagicMock(name="tensorflow.keras")
    tf.summary = mock.MagicMock(name="tensorflow.summary")
    tf.data = mock.MagicMock(name="tensorflow.data")
    tf.TensorSpec = mock.MagicMock(name="tensorflow.TensorSpec")
    return tf


def install() -> types.ModuleType:
    """Put the shim (and mocks for the absent I/O packages) into ``sys.modules``."""
    tf = build_module()
    sys.modules["tensorflow"] = tf
    sys.modules["tensorflow.nn"] = tf.nn
    sys.modules["tensorflow.math"] = tf.math
    sys.modules["tensorflow.keras"] = tf.keras
    for name in ("pycocotools", "pycocotools.coco", "skimage", "skimage.io"):
        if name not in sys.modules:
            sys.modules[name] = mock.MagicMock(name=name)
    return tf
