"""TEST INFRASTRUCTURE ONLY -- CPU (NumPy) restatement of the reference's SSD
box-geometry hot path.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this
module; the product (``ssdgeom``) never does and has no CPU fallback.

Every function cites the reference file:line (relative to /root/reference) whose
arithmetic it restates.  Pinning status:

* ``iou``, ``iou_n``, ``match_bbox``, ``apply_anchor_box``, ``build_prior_box``,
  ``ssd_loss``: PINNED -- checked against the reference's own golden values
  (tests/utils/test_bbox.py:10-17,35-44) and against outputs of the unmodified
  reference source executed in the build container (tests/golden/*.npz, made by
  oracle/make_golden.py; tests/test_oracle_vs_reference.py re-checks live whenever
  /root/reference is present).
* ``decode_bbox``: restates two inline lines (models/ssd_model.py:466-467);
  pinned by the encode->decode round trip only.
* ``score_head``: restates models/ssd_model.py:479-488 (softmax is TensorFlow's,
  un-vendored): parity unpinned by any reference test.
* ``nms_per_class``: PARITY UNPINNED -- the reference has no NMS at all
  (SURVEY.md section 0).  The spec is builder-defined from BASELINE.json
  (score>0.01, per-class top-k 200, IoU>0.45) using the reference's ``iou``
  formula (utils/bbox.py:13-25) in float32.

dtype discipline matters: the matcher's winners are decided by rounding noise
(hard part 1), so each side of the IoU is evaluated in the
dtype NumPy would use for the arrays the reference is called with (ground truth
float32, priors float64 on the training path) and promoted exactly as NumPy does.
"""
from __future__ import annotations

import math

import numpy as np

# ----------------------------------------------------------------------------------------
# A1  anchors
# ----------------------------------------------------------------------------------------

SSD300_SIZES = [(38, 38), (19, 19), (10, 10), (5, 5), (3, 3), (1, 1)]
SSD300_SK = [21, 45, 99, 153, 207, 261, 315]
SSD300_RATIOS = [[2], [2, 3], [2, 3], [2, 3], [2], [2]]


def build_prior_box(size_list=SSD300_SIZES, s_k_refer=SSD300_SK, aspect_ratio=SSD300_RATIOS,
                    input_size=300):
    """models/ssd_model.py:173-194 (``_build_prior_box``), with the two hard-coded tables
    (:176-177) and the input size (:184) lifted to arguments.  float64 [A,4] cxcywh,
    level-major, then y-outer / x-inner (:179), then the per-cell shape order of :185-192.
    """
    rows = []
    for lvl, (fh, fw) in enumerate(size_list):
        s = s_k_refer[lvl] / input_size
        s_next = s_k_refer[lvl + 1] / input_size
        s_prime = math.sqrt(s * s_next)
        shapes = [(s, s), (s_prime, s_prime)]
        for r in aspect_ratio[lvl]:
            shapes.append((s * math.sqrt(r), s / math.sqrt(r)))
            shapes.append((s / math.sqrt(r), s * math.sqrt(r)))
        for gy in range(fh):
            cy = (gy + 0.5) / fh
            for gx in range(fw):
                cx = (gx + 0.5) / fw
                for bw, bh in shapes:
                    rows.append((cx, cy, bw, bh))
    return np.array(rows, dtype=np.float64)


# ----------------------------------------------------------------------------------------
# A10 / A2  IoU
# ----------------------------------------------------------------------------------------

def iou(box_a, box_b, dtype=np.float32):
    """utils/bbox.py:6-25 (``iou``): element-wise IoU of cxcywh boxes, intersection extents
    clamped at 0.0 (:23), denominator + 1e-10 (:25).  Evaluated in ``dtype`` (float32 is what
    TensorFlow would use for float inputs)."""
    a = np.asarray(box_a, dtype=dtype)
    b = np.asarray(box_b, dtype=dtype)
    two = dtype(2)
    lo_x = np.maximum(a[..., 0] - a[..., 2] / two, b[..., 0] - b[..., 2] / two)
    lo_y = np.maximum(a[..., 1] - a[..., 3] / two, b[..., 1] - b[..., 3] / two)
    hi_x = np.minimum(a[..., 0] + a[..., 2] / two, b[..., 0] + b[..., 2] / two)
    hi_y = np.minimum(a[..., 1] + a[..., 3] / two, b[..., 1] + b[..., 3] / two)
    inter = np.maximum(dtype(0), hi_x - lo_x) * np.maximum(dtype(0), hi_y - lo_y)
    return inter / (a[..., 2] * a[..., 3] + b[..., 2] * b[..., 3] - inter + dtype(1e-10))


def iou_all_pairs(boxes, dtype=np.float32):
    """``iou`` (utils/bbox.py:6-25) of every pair of rows of ``boxes`` [..., m, 4] -> [..., m, m]: the same
    element-wise operations in the same order and dtype -- the per-box terms (corners, areas) are evaluated
    once per box instead of once per pair, which changes no bit (asserted in tests/test_oracle.py)."""
    b = np.asarray(boxes, dtype=dtype)
    two = dtype(2)
    x1, y1 = b[..., 0] - b[..., 2] / two, b[..., 1] - b[..., 3] / two
    x2, y2 = b[..., 0] + b[..., 2] / two, b[..., 1] + b[..., 3] / two
    area = b[..., 2] * b[..., 3]
    ex = np.minimum(x2[..., :, None], x2[..., None, :]) - np.maximum(x1[..., :, None], x1[..., None, :])
    np.maximum(dtype(0), ex, out=ex)
    ey = np.minimum(y2[..., :, None], y2[..., None, :]) - np.maximum(y1[..., :, None], y1[..., None, :])
    np.maximum(dtype(0), ey, out=ey)
    ex *= ey                                                              # inter
    den = area[..., :, None] + area[..., None, :]
    den -= ex
    den += dtype(1e-10)
    ex /= den
    return ex


def iou_matrix(gt_box, prior_box):
    """utils/bbox.py:28-41 (``iou_n``) applied to every (ground truth, prior) pair, i.e. what
    ``match_bbox`` builds at :53-58 with repeat/tile -- here by broadcasting, which performs the
    same element-wise operations in the same dtypes.  Extents clamp at 1e-10 (:39)."""
    g = np.asarray(gt_box)[:, None, :]
    p = np.asarray(prior_box)[None, :, :]
    g_area = g[..., 2] * g[..., 3]
    p_area = p[..., 2] * p[..., 3]
    lo_x = np.maximum(g[..., 0] - g[..., 2] / 2, p[..., 0] - p[..., 2] / 2)
    lo_y = np.maximum(g[..., 1] - g[..., 3] / 2, p[..., 1] - p[..., 3] / 2)
    hi_x = np.minimum(g[..., 0] + g[..., 2] / 2, p[..., 0] + p[..., 2] / 2)
    hi_y = np.minimum(g[..., 1] + g[..., 3] / 2, p[..., 1] + p[..., 3] / 2)
    inter = np.maximum(1e-10, hi_x - lo_x) * np.maximum(1e-10, hi_y - lo_y)
    return inter / (g_area + p_area - inter + 1e-10)


def iou_n(boxes_1, boxes_2):
    """utils/bbox.py:28-41: paired rows."""
    b1, b2 = np.asarray(boxes_1), np.asarray(boxes_2)
    a1 = b1[:, 2] * b1[:, 3]
    a2 = b2[:, 2] * b2[:, 3]
    lo_x = np.maximum(b1[:, 0] - b1[:, 2] / 2, b2[:, 0] - b2[:, 2] / 2)
    lo_y = np.maximum(b1[:, 1] - b1[:, 3] / 2, b2[:, 1] - b2[:, 3] / 2)
    hi_x = np.minimum(b1[:, 0] + b1[:, 2] / 2, b2[:, 0] + b2[:, 2] / 2)
    hi_y = np.minimum(b1[:, 1] + b1[:, 3] / 2, b2[:, 1] + b2[:, 3] / 2)
    inter = np.maximum(1e-10, hi_x - lo_x) * np.maximum(1e-10, hi_y - lo_y)
    return inter / (a1 + a2 - inter + 1e-10)


# ----------------------------------------------------------------------------------------
# A3  matcher
# ----------------------------------------------------------------------------------------

def _check_match_args(n_gt, n_prior, thresh):
    # utils/bbox.py:50-51
    assert n_gt <= n_prior, "more ground-truth boxes than priors"
    assert thresh > 0.0, "thresh must be positive"


def match_pairs_sweeps(iou_tp, thresh):
    """utils/bbox.py:60-79, the reference's own procedure: T rounds of whole-matrix arg-max
    with row+column knock-out on a scratch copy (:62-68), then arg-max sweeps of the matrix
    with the taken columns zeroed until the maximum is <= thresh (:71-79).  Returns the
    (gt, prior) pairs in the order the reference appends them."""
    n_gt, n_prior = iou_tp.shape
    left = iou_tp.copy()
    scratch = iou_tp.copy()
    pairs = []
    for _ in range(n_gt):
        flat = int(np.argmax(scratch))
        t, a = divmod(flat, n_prior)
        scratch[t, :] = 0.0
        scratch[:, a] = 0.0
        left[:, a] = 0.0
        pairs.append((t, a))
    while True:
        flat = int(np.argmax(left))
        t, a = divmod(flat, n_prior)
        if left[t, a] <= thresh:
            break
        pairs.append((t, a))
        left[:, a] = 0.0
    return pairs


def match_pairs_columnwise(iou_tp, thresh):
    """Same result as ``match_pairs_sweeps`` without the O(M) sweeps: after the T greedy
    rounds, every prior column that was not taken is positive iff its first-arg-max row
    exceeds ``thresh`` (utils/bbox.py:71-79 only ever zeroes whole columns, so columns are
    independent; the append order inside this phase does not matter to the scatter at
    :87-90 because its columns are distinct).  Used for sizes where the sweeps take minutes;
    equality with the sweeps is asserted in tests/test_oracle.py."""
    n_gt, n_prior = iou_tp.shape
    scratch = iou_tp.copy()
    taken = np.zeros(n_prior, dtype=bool)
    pairs = []
    for _ in range(n_gt):
        flat = int(np.argmax(scratch))
        t, a = divmod(flat, n_prior)
        scratch[t, :] = 0.0
        scratch[:, a] = 0.0
        taken[a] = True
        pairs.append((t, a))
    best_t = np.argmax(iou_tp, axis=0)
    best_v = iou_tp[best_t, np.arange(n_prior)]
    with np.errstate(invalid="ignore"):
        hit = ~(best_v <= thresh) & ~taken
    # order the second phase the way the reference would (descending value, first flat index)
    cols = np.nonzero(hit)[0]
    order = np.lexsort((cols, best_t[cols], -best_v[cols]))
    for a in cols[order]:
        pairs.append((int(best_t[a]), int(a)))
    return pairs


def scatter_pairs(pairs, gt_cls, gt_box, n_prior):
    """utils/bbox.py:84-90: later pairs overwrite earlier ones; untouched priors stay
    (cls 0, box 0, mask False)."""
    mask = np.zeros((n_prior,), dtype=bool)
    boxes = np.zeros((n_prior, 4), dtype=np.float32)
    labels = np.zeros((n_prior,), dtype=np.int32)
    for t, a in pairs:
        mask[a] = True
        boxes[a, :] = gt_box[t, :]
        labels[a] = int(gt_cls[t])
    return labels, boxes, mask


def match_bbox(cls, bbox, default_box, thresh=0.5, sweeps=True, return_pairs=False):
    """utils/bbox.py:44-91 (``match_bbox``)."""
    gt_cls, gt_box, priors = np.array(cls), np.array(bbox), np.array(default_box)
    n_gt, n_prior = gt_box.shape[0], priors.shape[0]
    _check_match_args(n_gt, n_prior, thresh)
    iou_tp = iou_matrix(gt_box, priors)
    pairs = (match_pairs_sweeps if sweeps else match_pairs_columnwise)(iou_tp, thresh)
    out = scatter_pairs(pairs, gt_cls, gt_box, n_prior)
    return out + (pairs,) if return_pairs else out


# ----------------------------------------------------------------------------------------
# A4 / A5 / A7  encode, target assignment, decode
# ----------------------------------------------------------------------------------------

def apply_anchor_box(origin_bbox, default_box):
    """utils/bbox.py:94-101 (``apply_anchor_box``): no variances; w/h clamped at 1e-5 in the
    operand's own dtype (:99)."""
    g, d = np.asarray(origin_bbox), np.asarray(default_box)
    assert g.shape == d.shape
    t_xy = (g[:, 0:2] - d[:, 0:2]) / d[:, 2:4]
    t_wh = np.log(np.maximum(g[:, 2:4], 1e-5) / np.maximum(d[:, 2:4], 1e-5))
    return np.concatenate([t_xy, t_wh], axis=-1)


def assign_encode(cls, bbox, prior_box, thresh=0.5, sweeps=True):
    """models/ssd_model.py:211-215 + the output_signature casts at :219-224: one image's
    (int32 [A], float32 [A,4], bool [A])."""
    labels, boxes, mask = match_bbox(cls, bbox, prior_box, thresh, sweeps=sweeps)
    loc = apply_anchor_box(boxes, prior_box).astype(np.float32)
    return labels.astype(np.int32), loc, mask.astype(bool)


def decode_bbox(loc, prior_box, scale=300.0, exp_dtype=np.float32):
    """models/ssd_model.py:466-467: ``xy = (t_xy*d_wh + d_xy)*scale``, ``wh = exp(t_wh)*d_wh*scale``.
    The reference holds the offsets in a float32 array, so ``np.exp`` runs in float32 while
    the products with the float64 priors run in float64 and the store rounds to float32.
    ``exp_dtype=np.float64`` gives the correctly-rounded variant the (builder-defined) NMS
    spec uses."""
    t = np.asarray(loc, dtype=np.float32)
    d = np.asarray(prior_box)
    out = np.empty(t.shape, dtype=np.float32)
    out[..., 0:2] = (t[..., 0:2] * d[..., 2:4] + d[..., 0:2]) * scale
    out[..., 2:4] = np.exp(t[..., 2:4].astype(exp_dtype)) * d[..., 2:4] * scale
    return out


# ----------------------------------------------------------------------------------------
# A6  multibox loss
# ----------------------------------------------------------------------------------------

def softmax_ce(logits, labels):
    """tf.nn.sparse_softmax_cross_entropy_with_logits as called at models/ssd_model.py:357,366
    (published formula; TensorFlow is un-vendored): float64 evaluation, float32 result."""
    x = np.asarray(logits).astype(np.float64)
    m = x.max(axis=-1, keepdims=True)
    lse = np.log(np.exp(x - m).sum(axis=-1))
    picked = np.take_along_axis(x - m, np.asarray(labels).astype(np.int64)[..., None], axis=-1)[..., 0]
    return (lse - picked).astype(np.float32)


def hard_negative_select(neg_ce, num_pos, ratio=3):
    """models/ssd_model.py:368-372: threshold = k-th largest of the flattened per-prior
    background CE (k = ratio*num_pos over the whole batch); every value >= threshold is kept,
    so ties can exceed ratio:1."""
    flat = np.asarray(neg_ce).reshape(-1)
    k = int(ratio) * int(num_pos)
    if k < 1 or k > flat.size:
        raise ValueError("hard-negative k=%d outside [1,%d]" % (k, flat.size))
    kth = np.partition(flat, flat.size - k)[flat.size - k]
    return kth, np.asarray(neg_ce) >= kth


def ssd_loss(y_true, y_pred, ratio=3, return_masks=False):
    """models/ssd_model.py:341-396 (``_ssd_loss``).  Background is the last class (:365), the
    localisation term is plain L1 (:384-386), each term is normalised by its own count."""
    gt_cls, gt_box, gt_mask = (np.asarray(v) for v in y_true)
    pred_box, pred_cls = (np.asarray(v) for v in y_pred)
    assert gt_cls.shape[0] == gt_box.shape[0] == gt_mask.shape[0] == pred_box.shape[0] == pred_cls.shape[0]
    assert gt_cls.shape[1] == pred_cls.shape[1]
    pos = gt_mask.astype(bool)
    n_pos = int(pos.sum())
    ce_gt = softmax_ce(pred_cls, gt_cls)
    loss_pos = float(np.sum(ce_gt[pos], dtype=np.float64) / n_pos) if n_pos else float("nan")
    bg = np.full(gt_cls.shape, pred_cls.shape[-1] - 1, dtype=np.int64)
    neg_ce = softmax_ce(pred_cls, bg) * (~pos).astype(np.float32)
    kth, neg = hard_negative_select(neg_ce, n_pos, ratio)
    assert not np.any(neg & pos)  # :375
    loss_neg = float(np.sum(neg_ce[neg], dtype=np.float64) / int(neg.sum()))
    l1 = np.abs(pred_box.astype(np.float32) - gt_box.astype(np.float32)).astype(np.float64).sum(axis=-1)
    loss_loc = float(np.sum(l1[pos]) / n_pos)
    info = {"cls loss pos": loss_pos, "cls loss neg": loss_neg, "loc loss": loss_loc}
    total = loss_loc + loss_pos + loss_neg
    if return_masks:
        return total, info, dict(neg_ce=neg_ce, neg_mask=neg, kth=kth, num_pos=n_pos, num_neg=int(neg.sum()))
    return total, info


def ssd_loss_grad(y_true, y_pred, ratio=3):
    """Analytic gradient of ``ssd_loss`` w.r.t. (pred_box, pred_cls) -- what ``tape.gradient``
    (models/ssd_model.py:248) back-propagates through :355-386; the masks and the mining
    threshold are piecewise constant.  float64 evaluation."""
    gt_cls, gt_box, gt_mask = (np.asarray(v) for v in y_true)
    pred_box, pred_cls = (np.asarray(v) for v in y_pred)
    _, _, aux = ssd_loss(y_true, y_pred, ratio, return_masks=True)
    pos = gt_mask.astype(bool)
    neg = aux["neg_mask"]
    x = pred_cls.astype(np.float64)
    p = np.exp(x - x.max(axis=-1, keepdims=True))
    p /= p.sum(axis=-1, keepdims=True)
    n_cls = pred_cls.shape[-1]
    w_pos = pos.astype(np.float64) / aux["num_pos"]
    w_neg = neg.astype(np.float64) / aux["num_neg"]
    g_cls = p * (w_pos + w_neg)[..., None]
    np.subtract.at(g_cls, tuple(np.nonzero(pos)) + (gt_cls[pos].astype(np.int64),), w_pos[pos])
    g_cls[..., n_cls - 1] -= w_neg
    g_box = np.sign(pred_box.astype(np.float32) - gt_box.astype(np.float32)).astype(np.float64) * w_pos[..., None]
    return g_box, g_cls


# ----------------------------------------------------------------------------------------
# A8  score head, A9 per-class NMS
# ----------------------------------------------------------------------------------------

def softmax(logits):
    """tf.nn.softmax at models/ssd_model.py:479 (published formula): float64 evaluation,
    float32 result."""
    x = np.asarray(logits).astype(np.float64)
    e = np.exp(x - x.max(axis=-1, keepdims=True))
    return (e / e.sum(axis=-1, keepdims=True)).astype(np.float32)


def score_head(pred_cls, thresh=0.5, probs=None):
    """models/ssd_model.py:479-488 with ``mask=None``: score = max foreground probability,
    keep = score>thresh and not background>thresh, cls = arg-max over all classes incl.
    background (first maximum)."""
    p = softmax(pred_cls) if probs is None else np.asarray(probs)
    score = p[..., :-1].max(axis=-1)
    keep = (score > thresh) & ~(p[..., -1] > thresh)
    return score, np.argmax(p, axis=-1), keep


def nms_single_class(scores, boxes, score_thresh=0.01, top_k=200, iou_thresh=0.45):
    """One class of one image.  ``scores`` float32 [A], ``boxes`` float32 [A,4] cxcywh.
    Candidates: score > score_thresh (strict).  Visit order: score descending, ties by
    lower prior index; only the first ``top_k`` are visited.  A visited candidate is kept
    unless a previously kept one has ``iou`` (utils/bbox.py:13-25, float32) > iou_thresh
    (strict).  Returns the kept prior indices in visit order."""
    s = np.asarray(scores, dtype=np.float32)
    cand = np.nonzero(s > np.float32(score_thresh))[0]
    if cand.size == 0:
        return np.zeros((0,), dtype=np.int32)
    order = np.lexsort((cand, -s[cand].astype(np.float64)))
    cand = cand[order][:top_k]
    b = np.asarray(boxes, dtype=np.float32)[cand]
    alive = np.ones(cand.size, dtype=bool)
    kept = []
    thr = np.float32(iou_thresh)
    for i in range(cand.size):
        if not alive[i]:
            continue
        kept.append(int(cand[i]))
        if i + 1 < cand.size:
            with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
                ov = iou(b[i], b[i + 1:], dtype=np.float32)
            alive[i + 1:] &= ~(ov > thr)
    return np.asarray(kept, dtype=np.int32)


def nms_per_class(probs, boxes, score_thresh=0.01, top_k=200, iou_thresh=0.45):
    """A9 for one image: ``probs`` float32 [A,C] (softmax output, background last),
    ``boxes`` float32 [A,4] decoded cxcywh.  Returns (kept int32 [C-1, top_k] padded with -1,
    counts int32 [C-1])."""
    n_fg = probs.shape[1] - 1
    kept = np.full((n_fg, top_k), -1, dtype=np.int32)
    counts = np.zeros((n_fg,), dtype=np.int32)
    for c in range(n_fg):
        k = nms_single_class(probs[:, c], boxes, score_thresh, top_k, iou_thresh)
        kept[c, :k.size] = k
        counts[c] = k.size
    return kept, counts


def nms_per_class_batched(probs, boxes, score_thresh=0.01, top_k=200, iou_thresh=0.45):
    """Same result as ``nms_per_class`` (asserted in tests/test_oracle.py), vectorised over the classes
    so that the CPU baseline of bench.py is not dominated by interpreter overhead: the visit lists of all
    classes are padded to one [C-1, m] array, the ``iou`` formula (utils/bbox.py:13-25, float32, the very
    same element-wise operations) is evaluated once for all [C-1, m, m] pairs, and the greedy visit runs
    as m steps over all classes at once."""
    probs = np.asarray(probs, dtype=np.float32)
    boxes = np.asarray(boxes, dtype=np.float32)
    n_fg = probs.shape[1] - 1
    kept = np.full((n_fg, top_k), -1, dtype=np.int32)
    counts = np.zeros((n_fg,), dtype=np.int32)
    lists = []
    for c in range(n_fg):
        s = probs[:, c]
        cand = np.nonzero(s > np.float32(score_thresh))[0]
        if cand.size:
            cand = cand[np.lexsort((cand, -s[cand].astype(np.float64)))][:top_k]
        lists.append(cand)
    m = max((l.size for l in lists), default=0)
    if m == 0:
        return kept, counts
    idx = np.zeros((n_fg, m), dtype=np.int64)
    live = np.zeros((n_fg, m), dtype=bool)
    for c, l in enumerate(lists):
        idx[c, :l.size] = l
        live[c, :l.size] = True
    b = boxes[idx]                                                     # [C-1, m, 4]
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        sup = iou_all_pairs(b) > np.float32(iou_thresh)                # [C-1, m(visited), m(later)]
    alive = live.copy()
    for i in range(m):
        k = alive[:, i]                     # classes whose i-th candidate is kept
        if i + 1 < m:
            alive[:, i + 1:] &= ~(sup[:, i, i + 1:] & k[:, None])
    for c in range(n_fg):
        k = idx[c, alive[c]].astype(np.int32)
        kept[c, :k.size] = k
        counts[c] = k.size
    return kept, counts


def detect(pred_cls, pred_box, prior_box, score_thresh=0.01, top_k=200, iou_thresh=0.45, batched=True):
    """Decode (scale 1.0) + softmax + per-class NMS for one image."""
    probs = softmax(pred_cls)
    boxes = decode_bbox(pred_box, prior_box, scale=1.0, exp_dtype=np.float64)
    kept, counts = (nms_per_class_batched if batched else nms_per_class)(probs, boxes, score_thresh, top_k, iou_thresh)
    return kept, counts, probs, boxes


# ---- input glue ------------------------------------------------------------
def coco_to_ssd_boxes(xywh, img_w, img_h):
    """One image: COCO pixel [x,y,w,h] -> relative cxcywh float32.  Follows
    data_loaders/coco/make_dataset.py:132 (centre = corner + size/2, in the array's own dtype), the float32
    TensorSpec of the generator (:140-142), and data_loaders/ssd/make_dataset.py:43-44 (in-place division of
    the float32 array by the integer [w,h,w,h])."""
    box = np.array(xywh).reshape(-1, 4)
    box[:, :2] += box[:, 2:] / 2
    box = box.astype(np.float32)
    scale = np.array([img_w, img_h, img_w, img_h])
    box /= scale
    return box


def normalize_image(image):
    """models/ssd_model.py:214 on the float32 image the loaders yield."""
    return (np.asarray(image, dtype=np.float32) - 0.5) * 2


# ---- generalised anchor tables: options the reference's rule does not have -------
def clip_priors(priors):
    """Component-wise clamp of cxcywh priors to [0,1] (the `clip` option of SSD variants)."""
    return np.clip(np.asarray(priors), 0.0, 1.0)


def apply_anchor_box_var(origin_bbox, default_box, variances):
    """utils/bbox.py:94-101 followed by the division by (v_xy, v_xy, v_wh, v_wh) of SSD variants; float64."""
    enc = apply_anchor_box(origin_bbox, default_box).astype(np.float64)
    return enc / np.array([variances[0], variances[0], variances[1], variances[1]], dtype=np.float64)


def decode_bbox_var(loc, default_box, scale, variances):
    """models/ssd_model.py:466-467 on offsets multiplied by the variances first (float32 product, like the kernel)."""
    v = np.array([variances[0], variances[0], variances[1], variances[1]], dtype=np.float32)
    return decode_bbox(np.asarray(loc, dtype=np.float32) * v, default_box, scale)
