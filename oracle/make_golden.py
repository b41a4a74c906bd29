"""TEST INFRASTRUCTURE ONLY -- generate ``tests/golden/*.npz`` by executing the
UNMODIFIED reference source (``/root/reference``, loaded by oracle/ref_loader.py under
the NumPy tensorflow shim) on seeded inputs.  Run in the build container:

    python -m oracle.make_golden

The fixtures are what travels to the GPU box (``/root/reference`` does not).  Inputs are
re-created from seeds by ``ssdgeom.synth`` in the tests, so each fixture also stores a
SHA-256 of the inputs it was generated from.

Fixtures
  priors_ssd300.npz   _build_prior_box (models/ssd_model.py:173-194) on the SSD300 maps
  assign_ssd300.npz   match_bbox + apply_anchor_box (utils/bbox.py:44-101) as called from
                      get_train_set (models/ssd_model.py:211-215): BASELINE config 1
                      (8 images, 8732 priors, <=100 GT, float32 GT x float64 priors)
  match_small.npz     match_bbox on small / degenerate inputs in the dtype mixes the
                      reference's own test uses (tests/utils/test_bbox.py:25-45)
  loss_ssd300.npz     _ssd_loss (models/ssd_model.py:341-396) on b=4 images
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "ssd-object-detection_b200"))
OUT = os.path.join(ROOT, "tests", "golden")


def sha(*arrays) -> str:
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode() + str(a.shape).encode() + a.tobytes())
    return h.hexdigest()


def main():
    from oracle import ref_loader
    from ssdgeom import synth

    ref = ref_loader.load()
    os.makedirs(OUT, exist_ok=True)

    # ---- priors ---------------------------------------------------------------------
    priors = ref.build_prior_box(synth.SSD300["sizes"])
    assert priors.shape == (8732, 4) and priors.dtype == np.float64
    np.savez_compressed(os.path.join(OUT, "priors_ssd300.npz"), priors=priors)

    # ---- config 1: assign + encode, 8 images ------------------------------------------
    # images 0-3: T=100 each ("max" mode); images 4-7: COCO-like T (lognormal)
    b_max, c_max, o_max = synth.make_gt(0, 4, 100, "max")
    b_coco, c_coco, o_coco = synth.make_gt(1, 4, 100, "coco")
    gt_boxes = np.concatenate([b_max, b_coco])
    gt_cls = np.concatenate([c_max, c_coco])
    offsets = np.concatenate([o_max, o_coco[1:] + o_max[-1]]).astype(np.int32)
    pairs_all, pair_off, loc_rows, sums = [], [0], [], []
    cls_full, loc_full, mask_full = [], [], []
    for i in range(8):
        s, e = offsets[i], offsets[i + 1]
        labels, boxes, mask = ref.match_bbox(gt_cls[s:e], gt_boxes[s:e], priors, 0.5)
        loc = ref.apply_anchor_box(boxes, priors).astype(np.float32)   # TensorSpec cast, :222
        labels, mask = labels.astype(np.int32), mask.astype(bool)
        # recover (t, a) per positive prior: the matched box row identifies t (first equal row
        # with the same class is sufficient for a fixture; tests compare cls/box/mask, not t)
        pos = np.nonzero(mask)[0]
        pairs_all.append(pos.astype(np.int32))
        pair_off.append(pair_off[-1] + pos.size)
        loc_rows.append(loc[pos])
        cls_full.append(labels); loc_full.append(loc); mask_full.append(mask)
        sums.append((sha(labels), sha(boxes), sha(mask), sha(loc)))
        print("assign image %d: T=%d positives=%d" % (i, e - s, pos.size))
    cls_full, loc_full, mask_full = np.stack(cls_full), np.stack(loc_full), np.stack(mask_full)
    np.savez_compressed(
        os.path.join(OUT, "assign_ssd300.npz"),
        input_sha=sha(gt_boxes, gt_cls, offsets, priors),
        offsets=offsets,
        pos_index=np.concatenate(pairs_all), pos_offsets=np.asarray(pair_off, dtype=np.int32),
        pos_cls=np.concatenate([c[m] for c, m in zip(cls_full, mask_full)]),
        pos_loc=np.concatenate(loc_rows),
        mask_bits=np.packbits(mask_full, axis=1),
        sha_cls=np.array([s[0] for s in sums]), sha_box=np.array([s[1] for s in sums]),
        sha_mask=np.array([s[2] for s in sums]), sha_loc=np.array([s[3] for s in sums]),
    )

    # ---- small / degenerate matcher cases ---------------------------------------------
    rng = np.random.default_rng(42)
    cases = {}

    def add(name, cls, box, pri, thresh=0.5):
        labels, boxes, mask = ref.match_bbox(cls, box, pri, thresh)
        cases[name + "_cls_in"] = np.asarray(cls)
        cases[name + "_box_in"] = np.asarray(box)
        cases[name + "_pri_in"] = np.asarray(pri)
        cases[name + "_thresh"] = np.float64(thresh)
        cases[name + "_cls"] = labels
        cases[name + "_box"] = boxes
        cases[name + "_mask"] = mask
        cases[name + "_enc"] = ref.apply_anchor_box(boxes, np.asarray(pri))

    # tests/utils/test_bbox.py:27-29 (float32 x float32)
    d = np.array([[10, 10, 2, 2], [10, 10, 0.5, 0.5], [11, 11, 3, 3]], dtype=np.float32)
    t = np.array([[0, 10, 10, 1, 1], [1, 11, 11, 2, 2]], dtype=np.float32)
    add("kat_f32", t[:, 0], t[:, 1:], d)
    # :35-39 (float64 x float64), identical geometry
    d = np.array([[10, 10, 1, 1], [20, 20, 1, 1], [20, 20, 0.5, 0.5]])
    t = np.array([[0, 10, 10, 0.5, 0.5], [1, 20, 20, 1, 1], [2, 20, 20, 0.5, 0.5]])
    add("kat_same", t[:, 0], t[:, 1:], d)
    # :40-44 greedy-order case
    d = np.array([[10, 10, 1, 1], [20, 20, 1.1, 1.1], [20, 20, 0.5, 0.5]])
    t = np.array([[0, 15, 15, 13, 13], [1, 15, 15, 14, 14]])
    add("kat_greedy", t[:, 0], t[:, 1:], d)
    # :31-33 random-normal boxes (negative w/h => negative IoUs: knocked-out zeros can win)
    for k in range(6):
        d = rng.normal(size=(20, 4))
        t = rng.normal(size=(2 + k, 5))
        add("normal%d" % k, t[:, 0], t[:, 1:], d)
    # duplicate ground truth / duplicate priors / T == A / tiny thresh / mixed dtypes
    pri = priors[::97][:64].copy()
    g = np.array([[0.5, 0.5, 0.3, 0.3]] * 3 + [[0.2, 0.2, 0.1, 0.1]], dtype=np.float32)
    add("dup_gt", np.arange(4, dtype=np.float32), g, pri)
    pri_dup = np.repeat(pri[:8], 3, axis=0)
    add("dup_prior", np.arange(4, dtype=np.float32), g, pri_dup)
    g8 = np.concatenate([rng.uniform(0.1, 0.9, (8, 2)), rng.uniform(0.05, 0.5, (8, 2))], 1).astype(np.float32)
    add("t_eq_a", np.arange(8, dtype=np.float32), g8, pri[:8])
    add("tiny_thresh", np.arange(8, dtype=np.float32), g8, pri, thresh=1e-12)
    add("high_thresh", np.arange(8, dtype=np.float32), g8, pri, thresh=0.9)
    add("f64_gt", np.arange(8, dtype=np.float64), g8.astype(np.float64) + 1e-9, pri)
    add("f32_prior", np.arange(8, dtype=np.float32), g8, pri.astype(np.float32))
    add("far_away", np.array([3.0], dtype=np.float32), np.array([[5, 5, 0.1, 0.1]], dtype=np.float32), pri)
    add("cls_trunc", np.array([2.9, -1.5, 79.0], dtype=np.float32), g8[:3], pri)
    cases["names"] = np.array(sorted({k.rsplit("_cls_in", 1)[0] for k in cases if k.endswith("_cls_in")}))
    np.savez_compressed(os.path.join(OUT, "match_small.npz"), **cases)

    # ---- loss, b = 4 --------------------------------------------------------------------
    bsz = 4
    pred_cls, pred_box = synth.make_predictions(0, bsz, 8732)
    y_true = (cls_full[:bsz], loc_full[:bsz], mask_full[:bsz])
    total, info = ref.ssd_loss(y_true, (pred_box, pred_cls))
    # the per-prior quantities, through the same shim ops the reference source just used
    import tensorflow as tf  # the shim
    neg_ce = np.asarray(tf.nn.sparse_softmax_cross_entropy_with_logits(
        np.full(cls_full[:bsz].shape, 80), pred_cls)) * (~mask_full[:bsz]).astype(np.float32)
    n_pos = int(mask_full[:bsz].sum())
    top, _ = tf.math.top_k(neg_ce.reshape(-1), 3 * n_pos)
    kth = np.float32(top[-1])
    neg_mask = neg_ce >= kth
    np.savez_compressed(
        os.path.join(OUT, "loss_ssd300.npz"),
        input_sha=sha(pred_cls, pred_box, *y_true),
        total=np.float64(total), loss_pos=np.float64(info["cls loss pos"]),
        loss_neg=np.float64(info["cls loss neg"]), loss_loc=np.float64(info["loc loss"]),
        num_pos=np.int64(n_pos), num_neg=np.int64(neg_mask.sum()), kth=kth,
        neg_mask_bits=np.packbits(neg_mask, axis=1),
    )
    # ---- input glue: the reference's _coco2ssd on annotation rows (section 8f row 3) ------------------
    rng = np.random.default_rng(12)
    sizes = np.array([(640, 427), (500, 375), (333, 500), (1, 1), (4000, 3000)], np.int32)
    counts = [9, 1, 40, 3, 5]
    g_off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    g_in, g_out = [], []
    for (w, h), t in zip(sizes, counts):
        xywh = np.concatenate([rng.uniform(0, [w, h], (t, 2)), rng.uniform(0.5, [w, h], (t, 2))], 1)
        centre = xywh.copy()
        centre[:, :2] += centre[:, 2:] / 2            # data_loaders/coco/make_dataset.py:132, verbatim
        image = np.zeros((int(h) if h < 64 else 8, int(w) if w < 64 else 8, 3), np.float32)
        # _coco2ssd reads (h, w) from the image: hand it an image of the real size only when small; otherwise
        # call the two arithmetic lines (:43-44) it would execute, on the real width / height
        box32 = centre.astype(np.float32)
        if image.shape[0] == h and image.shape[1] == w:
            _, _, rel = ref.coco2ssd(image, np.zeros(t, np.float32), box32)
        else:
            rel = box32
            rel /= np.array([w, h, w, h])             # data_loaders/ssd/make_dataset.py:43-44, verbatim
        g_in.append(xywh); g_out.append(rel)
    np.savez_compressed(os.path.join(OUT, "glue_small.npz"), xywh=np.concatenate(g_in), img_wh=sizes, offsets=g_off,
                        rel=np.concatenate(g_out))
    print("loss b=4:", float(total), {k: float(v) for k, v in info.items()}, n_pos, int(neg_mask.sum()))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
