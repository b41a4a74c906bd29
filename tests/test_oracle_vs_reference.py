"""CPU, build container only: run the UNMODIFIED reference source (oracle/ref_loader.py) next
to the oracle restatement on fresh random inputs.  Skipped where /root/reference is absent
(the GPU box); the committed fixtures in tests/golden cover that case."""
import numpy as np
import pytest

from oracle import ref_loader, ssd_oracle as O
from ssdgeom import synth

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference source not present")


@pytest.fixture(scope="module")
def ref():
    return ref_loader.load()


def test_reference_own_unit_tests_pass(ref):
    import subprocess, sys, os, textwrap
    code = textwrap.dedent("""
        import sys, unittest
        sys.path.insert(0, %r)
        from oracle import tf_shim; tf_shim.install()
        sys.path.insert(0, %r)
        import tests.utils.test_bbox as m
        r = unittest.TextTestRunner(verbosity=0).run(unittest.defaultTestLoader.loadTestsFromModule(m))
        sys.exit(0 if r.wasSuccessful() and r.testsRun == 3 else 1)
    """) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), ref_loader.REFERENCE_ROOT)
    # run outside this repo's tests/ package so `tests.utils` resolves to the reference's
    out = subprocess.run([sys.executable, "-c", code], cwd="/tmp", capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-2000:]


def test_priors(ref):
    assert np.array_equal(ref.build_prior_box(O.SSD300_SIZES), O.build_prior_box())


@pytest.mark.parametrize("seed,t,mode", [(11, 100, "max"), (12, 100, "coco"), (13, 37, "max")])
def test_assign_encode(ref, seed, t, mode):
    priors = O.build_prior_box()
    boxes, cls, off = synth.make_gt(seed, 2, t, mode)
    for i in range(2):
        s, e = off[i], off[i + 1]
        r_cls, r_box, r_mask = ref.match_bbox(cls[s:e], boxes[s:e], priors, 0.5)
        r_loc = ref.apply_anchor_box(r_box, priors)
        for sweeps in (True, False):
            o_cls, o_box, o_mask = O.match_bbox(cls[s:e], boxes[s:e], priors, 0.5, sweeps=sweeps)
            assert np.array_equal(o_cls, r_cls) and o_cls.dtype == r_cls.dtype
            assert np.array_equal(o_box, r_box) and o_box.dtype == r_box.dtype
            assert np.array_equal(o_mask, r_mask)
        assert np.array_equal(O.apply_anchor_box(r_box, priors), r_loc)


def test_match_degenerate(ref):
    rng = np.random.default_rng(99)
    for k in range(40):
        d = rng.normal(size=(int(rng.integers(3, 40)), 4))
        t = rng.normal(size=(int(rng.integers(1, min(6, d.shape[0]) + 1)), 5))
        if k % 3 == 0:
            d, t = d.astype(np.float32), t.astype(np.float32)
        thr = float(rng.choice([0.5, 0.05, 2.0]))
        with np.errstate(all="ignore"):
            want = ref.match_bbox(t[:, 0], t[:, 1:], d, thr)
            for sweeps in (True, False):
                got = O.match_bbox(t[:, 0], t[:, 1:], d, thr, sweeps=sweeps)
                for w, g in zip(want, got):
                    assert np.array_equal(w, g), (k, sweeps)


def test_loss(ref):
    priors = O.build_prior_box()
    boxes, cls, off = synth.make_gt(21, 3, 100, "coco")
    tgt = [O.assign_encode(cls[off[i]:off[i + 1]], boxes[off[i]:off[i + 1]], priors, sweeps=False) for i in range(3)]
    y_true = tuple(np.stack([t[k] for t in tgt]) for k in range(3))
    pred_cls, pred_box = synth.make_predictions(21, 3, 8732)
    r_total, r_info = ref.ssd_loss(y_true, (pred_box, pred_cls))
    o_total, o_info = O.ssd_loss(y_true, (pred_box, pred_cls))
    np.testing.assert_allclose(o_total, float(r_total), rtol=1e-7)
    for k in ("cls loss pos", "cls loss neg", "loc loss"):
        np.testing.assert_allclose(o_info[k], float(r_info[k]), rtol=1e-7)


def test_input_glue(ref):
    """data_loaders/coco/make_dataset.py:132 + data_loaders/ssd/make_dataset.py:37-46 against the restatement."""
    rng = np.random.default_rng(12)
    for (w, h, t) in [(640, 427, 9), (500, 375, 1), (333, 500, 40), (1, 1, 3)]:
        xywh = np.concatenate([rng.uniform(0, [w, h], (t, 2)), rng.uniform(0.5, [w, h], (t, 2))], 1)   # float64, as json.load gives
        cls = rng.integers(0, 80, t).astype(np.float32)
        image = rng.uniform(0, 1, (h, w, 3)).astype(np.float32)
        centre = xywh.copy()
        centre[:, :2] += centre[:, 2:] / 2                       # the generator's line 132, verbatim
        box32 = centre.astype(np.float32)                        # TensorSpec(..., tf.float32), :140-142
        r_image, r_cls, r_box = ref.coco2ssd(image, cls, box32)
        assert r_box.dtype == np.float32 and r_image.shape == (300, 300, 3)
        assert np.array_equal(r_box, O.coco_to_ssd_boxes(xywh, w, h))
    x = rng.uniform(0, 1, (4, 5, 3)).astype(np.float32)
    assert np.array_equal(O.normalize_image(x), (x - 0.5) * 2)
