"""CPU: pin the oracle restatement (oracle/ssd_oracle.py) against (i) the reference's own
known-answer tests (tests/utils/test_bbox.py in /root/reference) and (ii) the fixtures the
unmodified reference produced (tests/golden/*.npz, oracle/make_golden.py)."""
import hashlib
import os

import numpy as np
import pytest

from oracle import ssd_oracle as O
from ssdgeom import synth


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode() + str(a.shape).encode() + a.tobytes())
    return h.hexdigest()


# reference tests/utils/test_bbox.py:10-17
IOU_KAT = [
    ([10, 10, 2, 2], [10, 10, 2, 2], 1.0),
    ([10, 10, 1, 1], [20, 20, 1, 1], 0.0),
    ([10, 10, 2, 2], [10, 10, 4, 4], 0.25),
    ([10, 10, 0, 0], [20, 20, 0, 0], 0.0),
    ([10, 10, -1, -1], [10, 10, -1, -1], 0.0),
    ([10, 10, 2, 2], [11, 11, 2, 2], 1 / 7),
    ([10, 10, 6, 6], [13, 13, 2, 2], 1 / 39),
    ([10, -10, 1, 1], [10, -10, 1, 1], 1.0),
]


@pytest.mark.parametrize("a,b,want", IOU_KAT)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_iou_known_answers(a, b, want, dtype):
    assert abs(float(O.iou(a, b, dtype=dtype)) - want) < 5e-5   # places=4


def test_iou_n_probe_values():
    # reference tests/utils/test_bbox.py:19-23 prints these; SURVEY section 4 recorded them
    a = np.array([[10, 10, 2, 2], [10, 10, 1, 1], [10, 10, 2, 2]], dtype=np.float32)
    b = np.array([[10, 10, 2, 2], [20, 20, 1, 1], [10, 10, 4, 4]], dtype=np.float32)
    got = O.iou_n(a, b)
    assert got.dtype == np.float32
    np.testing.assert_allclose(got, [1.0, 5.0000002e-21, 0.25], rtol=1e-6)


def test_match_known_answers():
    # reference tests/utils/test_bbox.py:35-39
    d = np.array([[10, 10, 1, 1], [20, 20, 1, 1], [20, 20, 0.5, 0.5]])
    t = np.array([[0, 10, 10, 0.5, 0.5], [1, 20, 20, 1, 1], [2, 20, 20, 0.5, 0.5]])
    cls, loc, mask = O.match_bbox(t[:, 0], t[:, 1:], d)
    np.testing.assert_almost_equal(loc, t[:, 1:])
    # :40-44
    d = np.array([[10, 10, 1, 1], [20, 20, 1.1, 1.1], [20, 20, 0.5, 0.5]])
    t = np.array([[0, 15, 15, 13, 13], [1, 15, 15, 14, 14]])
    cls, loc, mask = O.match_bbox(t[:, 0], t[:, 1:], d)
    np.testing.assert_almost_equal(loc, np.array([[15, 15, 14, 14], [15, 15, 13, 13], [0, 0, 0, 0]]))
    assert cls.tolist() == [1, 0, 0] and mask.tolist() == [True, True, False]


def test_match_asserts():
    pri = O.build_prior_box()[:4]
    g = np.tile(np.array([[0.5, 0.5, 0.2, 0.2]], dtype=np.float32), (5, 1))
    with pytest.raises(AssertionError):
        O.match_bbox(np.zeros(5, np.float32), g, pri)          # T > A, utils/bbox.py:50
    with pytest.raises(AssertionError):
        O.match_bbox(np.zeros(2, np.float32), g[:2], pri, 0.0)  # thresh, :51


def test_priors_golden(golden_dir):
    want = np.load(os.path.join(golden_dir, "priors_ssd300.npz"))["priors"]
    got = O.build_prior_box()
    assert got.dtype == np.float64 and got.shape == (8732, 4)
    assert np.array_equal(got, want)            # bit-exact
    t = synth.SSD300
    assert np.array_equal(O.build_prior_box(t["sizes"], t["s_k_refer"], t["aspect_ratio"], t["input_size"]), want)
    # SURVEY section 8(a) probe rows
    np.testing.assert_allclose(got[0], [0.01315789, 0.01315789, 0.07, 0.07], rtol=1e-6)
    np.testing.assert_allclose(got[-1], [0.5, 0.5, 0.6151829, 1.2303658], rtol=1e-6)
    assert synth.num_priors(synth.SSD300) == 8732 and synth.num_priors(synth.SSD512) == 24564


def _config1_inputs():
    b_max, c_max, o_max = synth.make_gt(0, 4, 100, "max")
    b_coco, c_coco, o_coco = synth.make_gt(1, 4, 100, "coco")
    return (np.concatenate([b_max, b_coco]), np.concatenate([c_max, c_coco]),
            np.concatenate([o_max, o_coco[1:] + o_max[-1]]).astype(np.int32))


@pytest.mark.parametrize("sweeps", [True, False])
def test_assign_config1_golden(golden_dir, sweeps):
    """BASELINE config 1: 8 SSD300 images; oracle == unmodified reference, bit for bit."""
    g = np.load(os.path.join(golden_dir, "assign_ssd300.npz"))
    boxes, cls, offsets = _config1_inputs()
    priors = O.build_prior_box()
    assert sha(boxes, cls, offsets, priors) == str(g["input_sha"])
    for i in range(8):
        s, e = offsets[i], offsets[i + 1]
        lab, box, mask = O.match_bbox(cls[s:e], boxes[s:e], priors, 0.5, sweeps=sweeps)
        loc = O.apply_anchor_box(box, priors).astype(np.float32)
        assert sha(lab.astype(np.int32)) == str(g["sha_cls"][i])
        assert sha(box) == str(g["sha_box"][i])
        assert sha(mask.astype(bool)) == str(g["sha_mask"][i])
        assert sha(loc) == str(g["sha_loc"][i])
        ps, pe = g["pos_offsets"][i], g["pos_offsets"][i + 1]
        assert np.array_equal(np.nonzero(mask)[0], g["pos_index"][ps:pe])
        assert np.array_equal(loc[mask], g["pos_loc"][ps:pe])


def test_match_small_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "match_small.npz"))
    for name in g["names"]:
        for sweeps in (True, False):
            with np.errstate(all="ignore"):
                lab, box, mask = O.match_bbox(g[name + "_cls_in"], g[name + "_box_in"], g[name + "_pri_in"],
                                              float(g[name + "_thresh"]), sweeps=sweeps)
                enc = O.apply_anchor_box(box, g[name + "_pri_in"])
            assert np.array_equal(lab, g[name + "_cls"]), name
            assert np.array_equal(box, g[name + "_box"]), name
            assert np.array_equal(mask, g[name + "_mask"]), name
            assert np.array_equal(enc, g[name + "_enc"], equal_nan=True), name


def test_columnwise_equals_sweeps_random():
    rng = np.random.default_rng(7)
    pri = O.build_prior_box()[::13]
    for trial in range(20):
        t = int(rng.integers(1, 30))
        b, c, _ = synth.make_gt(100 + trial, 1, t, "max")
        if trial % 4 == 0:              # duplicates
            b[t // 2:] = b[: t - t // 2]
        thr = float(rng.choice([0.5, 0.3, 0.7, 1e-9]))
        m = O.iou_matrix(b, pri)
        assert O.match_pairs_sweeps(m, thr)[:t] == O.match_pairs_columnwise(m, thr)[:t]
        assert sorted(O.match_pairs_sweeps(m, thr)[t:]) == sorted(O.match_pairs_columnwise(m, thr)[t:])


def test_loss_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "loss_ssd300.npz"))
    boxes, cls, offsets = _config1_inputs()
    priors = O.build_prior_box()
    tgt = [O.assign_encode(cls[offsets[i]:offsets[i + 1]], boxes[offsets[i]:offsets[i + 1]], priors, sweeps=False)
           for i in range(4)]
    y_true = tuple(np.stack([t[k] for t in tgt]) for k in range(3))
    pred_cls, pred_box = synth.make_predictions(0, 4, 8732)
    assert sha(pred_cls, pred_box, *y_true) == str(g["input_sha"])
    total, info, aux = O.ssd_loss(y_true, (pred_box, pred_cls), return_masks=True)
    assert aux["num_pos"] == int(g["num_pos"]) and aux["num_neg"] == int(g["num_neg"])
    assert np.float32(aux["kth"]) == g["kth"]
    assert np.array_equal(np.packbits(aux["neg_mask"], axis=1), g["neg_mask_bits"])
    np.testing.assert_allclose(total, float(g["total"]), rtol=1e-6)
    np.testing.assert_allclose(info["cls loss pos"], float(g["loss_pos"]), rtol=1e-6)
    np.testing.assert_allclose(info["cls loss neg"], float(g["loss_neg"]), rtol=1e-6)
    np.testing.assert_allclose(info["loc loss"], float(g["loss_loc"]), rtol=1e-6)


def test_loss_guards():
    y_true = (np.zeros((1, 8), np.int32), np.zeros((1, 8, 4), np.float32), np.zeros((1, 8), bool))
    y_pred = (np.zeros((1, 8, 4), np.float32), np.zeros((1, 8, 5), np.float32))
    with pytest.raises(ValueError):
        O.ssd_loss(y_true, y_pred)            # num_pos == 0 -> k = 0 (reference: index error)
    y_true[2][0, :4] = True
    with pytest.raises(ValueError):
        O.ssd_loss(y_true, y_pred)            # 3*4 > 8 (reference: top_k error)


def test_encode_decode_roundtrip():
    priors = O.build_prior_box()
    rng = np.random.default_rng(3)
    g = np.concatenate([rng.uniform(0, 1, (8732, 2)), np.exp(rng.uniform(np.log(0.02), np.log(0.9), (8732, 2)))],
                       1).astype(np.float32)
    enc = O.apply_anchor_box(g, priors).astype(np.float32)
    for ed in (np.float32, np.float64):
        dec = O.decode_bbox(enc, priors, scale=1.0, exp_dtype=ed)
        np.testing.assert_allclose(dec, g, rtol=2e-5, atol=2e-6)


def test_loss_grad_matches_finite_difference():
    rng = np.random.default_rng(5)
    b, a, c = 2, 40, 6
    gt_cls = rng.integers(0, c - 1, (b, a)).astype(np.int32)
    gt_mask = rng.uniform(size=(b, a)) < 0.1
    gt_mask[0, 0] = True
    gt_box = rng.normal(size=(b, a, 4)).astype(np.float32)
    pred_box = rng.normal(size=(b, a, 4)).astype(np.float32)
    pred_cls = rng.normal(size=(b, a, c)).astype(np.float32)
    g_box, g_cls = O.ssd_loss_grad((gt_cls, gt_box, gt_mask), (pred_box, pred_cls))
    # the shim CE is float32-rounded, so finite differences need a float64 re-evaluation
    def f64_loss(pc):
        x = pc.astype(np.float64)
        lse = np.log(np.exp(x - x.max(-1, keepdims=True)).sum(-1)) + x.max(-1)
        _, _, aux = O.ssd_loss((gt_cls, gt_box, gt_mask), (pred_box, pred_cls), return_masks=True)
        ce_gt = lse - np.take_along_axis(x, gt_cls[..., None].astype(np.int64), -1)[..., 0]
        ce_bg = lse - x[..., -1]
        return ce_gt[gt_mask].sum() / aux["num_pos"] + ce_bg[aux["neg_mask"]].sum() / aux["num_neg"]
    base = pred_cls.astype(np.float64)
    for idx in [(0, 0, 1), (1, 5, c - 1), (0, 7, 2)]:
        h = 1e-6
        up, dn = base.copy(), base.copy()
        up[idx] += h
        dn[idx] -= h
        fd = (f64_loss(up) - f64_loss(dn)) / (2 * h)
        assert abs(fd - g_cls[idx]) < 1e-6


def test_nms_single_class_basics():
    boxes = np.array([[0.5, 0.5, 0.2, 0.2], [0.5, 0.5, 0.21, 0.2], [0.1, 0.1, 0.1, 0.1],
                      [0.5, 0.52, 0.2, 0.2], [0.9, 0.9, 0.1, 0.1]], dtype=np.float32)
    scores = np.array([0.9, 0.8, 0.7, 0.005, 0.9], dtype=np.float32)
    kept = O.nms_single_class(scores, boxes)
    assert kept.tolist() == [0, 4, 2]          # tie 0.9: lower index first; 1 suppressed; 3 below thresh
    assert O.nms_single_class(scores, boxes, top_k=2).tolist() == [0, 4]
    assert O.nms_single_class(np.zeros(5, np.float32), boxes).size == 0


def test_nms_against_torchvision_sanity():
    tv = pytest.importorskip("torchvision")
    import torch
    rng = np.random.default_rng(11)
    n = 300
    b = np.concatenate([rng.uniform(0.2, 0.8, (n, 2)), rng.uniform(0.05, 0.3, (n, 2))], 1).astype(np.float32)
    s = rng.uniform(0.02, 1, n).astype(np.float32)
    kept = O.nms_single_class(s, b, top_k=n)
    xyxy = torch.tensor(np.concatenate([b[:, :2] - b[:, 2:] / 2, b[:, :2] + b[:, 2:] / 2], 1))
    ref = tv.ops.nms(xyxy, torch.tensor(s), 0.45).numpy()
    assert kept.tolist() == ref.tolist()


def test_input_glue_golden(golden_dir):
    """coco xywh pixels -> relative cxcywh: the reference's own outputs (oracle/make_golden.py)."""
    g = np.load(os.path.join(golden_dir, "glue_small.npz"))
    off = g["offsets"]
    for i, (w, h) in enumerate(g["img_wh"]):
        got = O.coco_to_ssd_boxes(g["xywh"][off[i]:off[i + 1]], int(w), int(h))
        assert got.dtype == np.float32 and np.array_equal(got, g["rel"][off[i]:off[i + 1]])
    x = np.linspace(-1, 2, 101, dtype=np.float32)
    assert np.array_equal(O.normalize_image(x), (x - np.float32(0.5)) * np.float32(2))


# ---- the pruning rules of the CUDA kernels, restated in float32 NumPy: necessary conditions only -------------
def _nms_join_candidates(boxes, thr):
    """The candidate pairs nms_kernel's interval join keeps (detect.cu: shrunk-extent slab intervals per axis plus
    width / height class neighbourhoods), restated with the kernel's float32 arithmetic."""
    f = np.float32
    b = boxes.astype(f)
    hw, hh = b[:, 2] * f(0.5), b[:, 3] * f(0.5)
    x1, y1, x2, y2 = b[:, 0] - hw, b[:, 1] - hh, b[:, 0] + hw, b[:, 1] + hh
    area = b[:, 2] * b[:, 3]
    q = f(thr) / (f(1) + f(thr))
    pr = (x2 - x1) * (y2 - y1)
    inexact = bool(np.any(~((area >= f(0.999) * pr) & (area <= f(1.001) * pr))))
    tq = f(0) if inexact else min(f(0.98) * q, f(0.49))
    nslab = 32

    def slab(v, o, sc):
        return np.clip(np.nan_to_num((v - o) * sc, nan=0.0), 0, nslab - 1).astype(np.int64)

    dx0, dy0 = x1.min(), y1.min()
    dsx = f(nslab) / (x2.max() - dx0) if x2.max() > dx0 else f(0)
    dsy = f(nslab) / (y2.max() - dy0) if y2.max() > dy0 else f(0)
    sx = tq * (x2 - x1) - f(1e-6) * (np.abs(x1) + np.abs(x2))
    sy = tq * (y2 - y1) - f(1e-6) * (np.abs(y1) + np.abs(y2))
    ax, bx = slab(x1 + sx, dx0, dsx), None
    bx = np.maximum(slab(x2 - sx, dx0, dsx), ax)
    ay = slab(y1 + sy, dy0, dsy)
    by = np.maximum(slab(y2 - sy, dy0, dsy), ay)
    tp = f(0.99) * f(thr) / (f(1) + f(0.001) * f(thr))
    if (not inexact) and tp < f(0.98):
        inv_l = f(-1) / np.log2(tp)
        lwx = np.log2(x2.max() - dx0) if x2.max() > dx0 else f(0)
        lwy = np.log2(y2.max() - dy0) if y2.max() > dy0 else f(0)
        cw = np.clip((lwx - np.log2(x2 - x1)) * inv_l, 0, 15).astype(np.int64)
        ch = np.clip((lwy - np.log2(y2 - y1)) * inv_l, 0, 15).astype(np.int64)
    else:
        cw = ch = np.zeros(len(b), np.int64)
    i, j = np.tril_indices(len(b), -1)
    keep = ((ax[j] <= bx[i]) & (bx[j] >= ax[i]) & (ay[j] <= by[i]) & (by[j] >= ay[i]) &
            (np.abs(cw[i] - cw[j]) <= 1) & (np.abs(ch[i] - ch[j]) <= 1))
    return i, j, keep


@pytest.mark.parametrize("thr", [0.05, 0.3, 0.45, 0.6, 0.9])
def test_nms_join_never_drops_a_suppressing_pair(thr):
    """Every pair the formula (utils/bbox.py:13-25, float32) puts above the threshold survives the join."""
    rng = np.random.default_rng(int(thr * 100))
    for scale in (0.05, 0.3, 1.0):
        n = 200
        c = rng.uniform(0, 1, (n, 2))
        wh = np.exp(rng.uniform(np.log(0.01), np.log(0.9), (n, 2))) * scale
        boxes = np.concatenate([c, wh], 1).astype(np.float32)
        dup = boxes[1::9][:len(boxes[::9])].copy()                              # near duplicates: real suppression
        dup[:, 2:] *= np.float32(1.01)
        boxes[::9] = dup
        i, j, keep = _nms_join_candidates(boxes, thr)
        iou = O.iou(boxes[i], boxes[j])
        hot = iou > np.float32(thr)
        assert hot.sum() > 0
        assert not np.any(hot & ~keep)
        assert keep.sum() < 0.5 * len(i)                                        # and it does prune


def test_filter_partial_sum_prefilter_is_necessary():
    """filter_kernel's pre-filter: with the background as exponent reference and the bound refreshed every 16 classes
    from the partial sum, e_c > 0.999 thr * max(1, partial) never rejects a class whose softmax score exceeds thr."""
    f = np.float32
    for bias, seed in ((7.0, 0), (0.0, 1), (3.0, 2)):
        x, _ = synth.make_predictions(seed, 1, 4096, 81, bg_bias=bias)
        x = x[0]
        p = O.softmax(x)
        e = np.exp2((x - x[:, -1:]) * f(1.4426950408889634)).astype(f)
        part = e[:, -1].copy()                                                  # the background opens the sum
        passed = np.zeros((x.shape[0], 80), bool)
        pre0 = f(0.01) * f(0.999)
        pre = np.full(x.shape[0], pre0, f)
        for c0 in range(0, 80, 16):
            blk = e[:, c0:c0 + 16]
            passed[:, c0:c0 + 16] = blk > pre[:, None]
            part = part + blk.sum(1, dtype=f)
            pre = np.maximum(pre0, pre0 * part)
        must = p[:, :80] > 0.01 * (1 + 1e-5)
        assert not np.any(must & ~passed)
        assert passed.sum() < 2.2 * max(must.sum(), 1) or bias == 0.0


# ---- the vectorised NMS restatement bench.py's CPU arm uses ------------------------------------------------------
def test_iou_all_pairs_is_the_iou_formula_bit_for_bit():
    rng = np.random.default_rng(3)
    b = rng.normal(size=(4, 50, 4)).astype(np.float32)          # includes negative / tiny sizes
    b[0, :5, 2:] = 0.0
    with np.errstate(all="ignore"):
        want = O.iou(b[:, :, None, :], b[:, None, :, :], dtype=np.float32)
        got = O.iou_all_pairs(b)
    assert got.dtype == np.float32 and np.array_equal(got, want, equal_nan=True)


@pytest.mark.parametrize("bias,thr,topk,iou_thr", [(7.0, 0.01, 200, 0.45), (0.0, 0.01, 200, 0.45), (5.0, 0.02, 50, 0.3),
                                                    (3.0, 0.5, 10, 0.6)])
def test_nms_batched_equals_per_class(bias, thr, topk, iou_thr):
    priors = O.build_prior_box()
    pred_cls, pred_box = synth.make_predictions(77, 1, priors.shape[0], bg_bias=bias)
    probs = O.softmax(pred_cls[0])
    boxes = O.decode_bbox(pred_box[0] * np.float32(0.4), priors, scale=1.0, exp_dtype=np.float64)
    k0, c0 = O.nms_per_class(probs, boxes, thr, topk, iou_thr)
    k1, c1 = O.nms_per_class_batched(probs, boxes, thr, topk, iou_thr)
    assert np.array_equal(c0, c1) and np.array_equal(k0, k1)
    assert c0.sum() > 0 or thr >= 0.5
