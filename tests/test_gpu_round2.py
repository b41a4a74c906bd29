"""GPU (B200): the boundary additions of round 2, called through the C ABI --
the library's own collective (ssdg_comm_*), gradients under cross-shard mining (stage 4), the matcher's status word
in the pipeline, the prior-index content check, class-id range check, the autograd bridge, the prefetching iterator,
and tighter evidence for the (parity-unpinned) NMS."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import ssd_oracle as O          # noqa: E402  (the checker)
from ssdgeom import synth                   # noqa: E402
from ssdgeom import _native as N, ops, device as D, parallel   # noqa: E402
from ssdgeom.models import ssd_model as M   # noqa: E402
from ssdgeom.pipeline import HotPath        # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def priors300():
    return M.build_prior_box(synth.SSD300["sizes"])


def _targets(boxes, cls, off, priors, n):
    out = ops.match_encode(boxes, cls, off, priors, n, int(np.diff(off).max()), 0.5)
    return out["cls"].to_host(), out["loc"].to_host(), out["mask"].to_host()


# ---- ssdg_comm_* ---------------------------------------------------------------------------------------------
def test_comm_single_rank_roundtrip():
    from ssdgeom import comm
    assert comm.available() and comm.nccl_version() >= 20000
    cm = comm.Comm(comm.unique_id(), 1, 0)
    for dt in (np.float64, np.float32, np.int32, np.int64):
        x = np.arange(1, 41).astype(dt)
        d = D.to_device(x)
        cm.allreduce(d)
        assert np.array_equal(d.to_host(), x)
    a, b = D.to_device(np.ones(2048, np.int32)), D.to_device(np.full(3, 2.5))
    cm.allreduce_multi([a, b])
    assert a.to_host().sum() == 2048 and np.array_equal(b.to_host(), np.full(3, 2.5))
    cm.close()


_TWO_RANK = r"""
import os, sys, numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "ssd-object-detection_b200"))
rank, world, port = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
from ssdgeom import _native as N, device as D, comm, ops, synth, parallel
from ssdgeom.models import ssd_model as M
N.check(N.lib().ssdg_set_device(rank), "set_device")
uid = comm.tcp_broadcast(comm.unique_id() if rank == 0 else None, world, rank, "127.0.0.1", port)
cm = comm.Comm(uid, world, rank)
st = D.Stream()
x = D.to_device(np.arange(8, dtype=np.float64) + rank)
cm.allreduce(x, st); st.sync()
assert np.array_equal(x.to_host(), 2 * np.arange(8) + 1.0), x.to_host()
# exact batch-global mining of a 6-image batch over the two devices, exchanged by ssdg_comm (one NCCL group per stage),
# with gradients: every rank must reproduce its slice of the single-device result bit for bit (mask) / to rounding
priors = M.build_prior_box(synth.SSD300["sizes"])
batch = 6
boxes, cls, off = synth.make_gt(31, batch, 100, "coco")
t = ops.match_encode(boxes, cls, off, priors, batch, int(np.diff(off).max()), 0.5)
y = (t["cls"].to_host(), t["loc"].to_host(), t["mask"].to_host())
pc, pb = synth.make_predictions(31, batch, 8732)
full = ops.multibox_loss(y[0], y[1], y[2], pb, pc, want_neg_mask=True, want_grad=True)
want = ops.loss_result_to_host(full["result"])
lo, hi = parallel.shard_range(batch, world, rank)
sl = ops.StagedLoss(y[0][lo:hi], y[1][lo:hi], y[2][lo:hi], pb[lo:hi], pc[lo:hi], global_priors=batch * 8732,
                    want_neg_mask=True, want_grad=True, stream=st)
for stage in range(sl.n_stages):
    sl.run(stage)
    bufs = sl.exchange(stage)
    if bufs:
        cm.allreduce_multi(bufs, st)
total, info = sl.finish()
assert info["num_pos"] == want["num_pos"] and info["num_neg"] == want["num_neg"] and info["kth"] == want["kth"]
np.testing.assert_allclose(total, want["total"], rtol=1e-9)
assert np.array_equal(sl.out["neg_mask"].to_host(), full["neg_mask"].to_host()[lo:hi])
np.testing.assert_allclose(sl.out["grad_cls"].to_host(), full["grad_cls"].to_host()[lo:hi], rtol=1e-6, atol=1e-12)
np.testing.assert_array_equal(sl.out["grad_box"].to_host(), full["grad_box"].to_host()[lo:hi])
cm.close()
print("rank", rank, "ok")
"""


def test_comm_two_ranks_allreduce_and_global_mining_with_gradients():
    if N.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = [subprocess.Popen([sys.executable, "-c", _TWO_RANK, ROOT, str(r), "2", str(port)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and "rank %d ok" % r in o, o[-3000:]


# ---- gradients under cross-shard mining (stage 3 scaled mined negatives by the LOCAL count) ------------
@pytest.mark.parametrize("shards", [2, 3])
def test_staged_loss_gradients_equal_single_device(shards, priors300):
    batch = 6
    boxes, cls, off = synth.make_gt(41, batch, 100, "coco")
    y_true = _targets(boxes, cls, off, priors300, batch)
    pred_cls, pred_box = synth.make_predictions(41, batch, 8732)
    full = ops.multibox_loss(y_true[0], y_true[1], y_true[2], pred_box, pred_cls, want_grad=True)
    want = ops.loss_result_to_host(full["result"])
    w_box, w_cls = full["grad_box"].to_host(), full["grad_cls"].to_host()
    o_box, o_cls = O.ssd_loss_grad((y_true[0], y_true[1], y_true[2].astype(bool)), (pred_box, pred_cls))
    np.testing.assert_allclose(w_cls, o_cls, rtol=2e-5, atol=1e-9)
    spans = [parallel.shard_range(batch, shards, r) for r in range(shards)]
    staged = [ops.StagedLoss(*(v[lo:hi] for v in y_true), pred_box[lo:hi], pred_cls[lo:hi], global_priors=batch * 8732,
                             want_grad=True) for lo, hi in spans]
    assert staged[0].n_stages == 5
    for stage in range(5):
        for s in staged:
            s.run(stage)
        for bufs in zip(*(s.exchange(stage) for s in staged)):      # the all-reduce, emulated on the host
            tot = sum(b.to_host().astype(np.float64 if b.dtype == np.float64 else np.int64) for b in bufs)
            for b in bufs:
                b.copy_from_host(tot.astype(b.dtype))
            D.sync()
    for s, (lo, hi) in zip(staged, spans):
        total, info = s.finish()
        assert info["num_neg"] == want["num_neg"] and info["num_pos"] == want["num_pos"]
        # same weights 1/num_pos, 1/num_neg (global) and the same softmax: equal to float rounding of 1/n
        np.testing.assert_allclose(s.out["grad_cls"].to_host(), w_cls[lo:hi], rtol=1e-6, atol=1e-12)
        np.testing.assert_array_equal(s.out["grad_box"].to_host(), w_box[lo:hi])


# ---- matcher status in the pipeline ------------------------------------------------------------------
def test_hotpath_refuses_batches_it_cannot_match(priors300):
    b = 2
    boxes, cls, off = synth.make_gt(3, b, 12, "max")
    pc, pb = synth.make_predictions(3, b, 8732)
    out = (np.zeros(16), np.zeros((b, 80, 200), np.int32), np.zeros((b, 80), np.int32))
    hp = HotPath(synth.SSD300, batch=b, max_gt=8, total_gt=boxes.shape[0])       # built for 8 GT, fed 12
    with pytest.raises(ValueError):
        hp.step_host(boxes, cls, off, pc, pb, *out)
    # the device-side guard behind it: the C entry point raises the status bit and the host layer turns it into an error
    res = ops.match_encode(boxes, cls, off, priors300, b, 8, 0.5)
    st = ops.match_status(res)
    assert st & ops.MATCH_STATUS_TOO_MANY_GT
    with pytest.raises(ValueError):
        ops.raise_for_match_status(st)
    assert not res["mask"].to_host().any()          # documented: such images are written as all-unmatched
    # a correctly sized pipeline reports a clean status through download()
    hp = HotPath(synth.SSD300, batch=b, max_gt=12, total_gt=boxes.shape[0])
    hp.step_host(boxes, cls, off, pc, pb, *out)
    assert int(hp._status_host.array[1]) & 11 == 0 and out[0][7] == 0


def test_prior_index_content_check_and_destroy():
    pri = ops.prior_boxes(synth.SSD300["sizes"], synth.SSD300["s_k_refer"], synth.SSD300["aspect_ratio"], 300)
    idx = ops.prior_index(pri)
    boxes, cls, off = synth.make_gt(8, 2, 10, "max")
    good = ops.match_encode(boxes, cls, off, pri, 2, 10, 0.5)
    assert ops.match_status(good) & ops.MATCH_STATUS_STALE_INDEX == 0
    want = (good["cls"].to_host(), good["mask"].to_host())
    # priors edited in place after the index was built (the un-clipped SSD300 table has boxes beyond [0,1])
    N.check(N.lib().ssdg_priors_clip(pri.ptr, N.F64, pri.shape[0], None), "clip")
    stale = ops.match_encode(boxes, cls, off, pri, 2, 10, 0.5)
    assert ops.match_status(stale) & ops.MATCH_STATUS_STALE_INDEX
    with pytest.raises(N.SsdgeomError):
        ops.raise_for_match_status(ops.match_status(stale))
    # without the index the edited priors are simply the priors
    del pri._ssdg_index
    plain = ops.match_encode(boxes, cls, off, pri, 2, 10, 0.5)
    assert ops.match_status(plain) & ops.MATCH_STATUS_STALE_INDEX == 0
    w = O.match_bbox(cls[off[0]:off[1]], boxes[off[0]:off[1]], pri.to_host(), 0.5, sweeps=False)
    assert np.array_equal(plain["cls"].to_host()[0], w[0]) and np.array_equal(plain["mask"].to_host()[0].astype(bool), w[2])
    assert want[0].shape == plain["cls"].shape
    # a destroyed index is refused
    ptr = idx.ptr
    assert N.lib().ssdg_prior_index_destroy(ptr) == N.OK
    assert N.lib().ssdg_prior_index_destroy(ptr) == N.ERR_ARG
    idx._built = False
    with pytest.raises(ValueError):
        ops.match_encode(boxes, cls, off, pri, 2, 10, 0.5, index=idx)


def test_loss_flags_class_ids_out_of_range(priors300):
    """tf.nn.sparse_softmax_cross_entropy_with_logits (models/ssd_model.py:357) raises on a label outside [0, C)."""
    boxes, cls, off = synth.make_gt(6, 2, 20, "max")
    y = list(_targets(boxes, cls, off, priors300, 2))
    pc, pb = synth.make_predictions(6, 2, 8732)
    pos = np.argwhere(y[2])
    y[0] = y[0].copy()
    y[0][tuple(pos[0])] = 81
    for kw in ({}, {"row_stats": True}):
        if kw:
            det = ops.detect(pc, pb, priors300, want_row_stats=True)
            kw = {"row_stats": (det["row_ml"], det["row_negbg"])}
        r = ops.multibox_loss(y[0], y[1], y[2], pb, pc, **kw)["result"].to_host()
        assert int(r[7]) == N.ERR_LABEL_RANGE and r[12] == 1 and np.isnan(r[0])
        with pytest.raises(ValueError):
            ops.loss_result_to_host(ops.multibox_loss(y[0], y[1], y[2], pb, pc, **kw)["result"])
    # labels of NON-positive priors are never looked at (the reference multiplies their CE by the mask)
    y[0][tuple(pos[0])] = 3
    y[0][~y[2].astype(bool)] = 999
    assert ops.multibox_loss(y[0], y[1], y[2], pb, pc)["result"].to_host()[7] == 0


# ---- A9: exact account of the lists that differ from the float64 oracle, torchvision cross-check ------------------
def _explain_difference(kept_g, kept_o, probs64, boxes, thr_iou, c):
    """A kept list may differ from the float64 oracle's only through a near-tie the float32 scores resolve the
    other way: two candidates whose float64 scores agree to 2e-5 (visit order / top-k cut), or a pair whose IoU is
    within 1e-5 of the threshold.  Returns the relative score gap at the first divergence."""
    n = min(len(kept_g), len(kept_o))
    first = next((i for i in range(n) if kept_g[i] != kept_o[i]), n)
    a1 = kept_g[first] if first < len(kept_g) else kept_o[first]
    a2 = kept_o[first] if first < len(kept_o) else kept_g[first]
    s1, s2 = probs64[a1, c], probs64[a2, c]
    return abs(s1 - s2) / max(s1, s2)


@pytest.mark.parametrize("bias,images", [(7.0, 4), (0.0, 1)])
def test_detect_end_to_end_differences_are_counted_and_explained(bias, images, priors300, capsys):
    pred_cls, pred_box = synth.make_predictions(131, images, 8732, bg_bias=bias)
    kept, count = M.detect(pred_cls, pred_box, priors300)
    differ, gaps, lists = 0, [], 0
    for i in range(images):
        w_kept, w_count, w_probs, w_boxes = O.detect(pred_cls[i], pred_box[i], priors300)
        x = pred_cls[i].astype(np.float64)
        p64 = np.exp(x - x.max(-1, keepdims=True))
        p64 /= p64.sum(-1, keepdims=True)
        for c in range(kept.shape[1]):
            lists += 1
            g, o = kept[i, c, :count[i, c]], w_kept[c, :w_count[c]]
            if len(g) == len(o) and np.array_equal(g, o):
                continue
            differ += 1
            gaps.append(_explain_difference(list(g), list(o), p64, w_boxes, 0.45, c))
    with capsys.disabled():
        print("\n[A9] bias %.0f: %d of %d kept lists differ from the float64 oracle end to end; relative score gaps at "
              "the first divergence: %s" % (bias, differ, lists, ", ".join("%.1e" % g for g in gaps) or "-"))
    assert differ <= max(2, lists // 100)
    assert all(g < 2e-5 for g in gaps), gaps      # every difference is a near-tie inside the 1e-5 score contract


def test_nms_single_class_against_torchvision_on_the_cuda_path():
    """Cross-check of the CUDA suppression itself (not of the oracle): well-separated scores, boxes away from the
    threshold, one class -> torchvision.ops.nms (xyxy, no 1e-10 term) must keep the same set in the same order."""
    tv = pytest.importorskip("torchvision")
    import torch
    rng = np.random.default_rng(5)
    a = 600
    centres = rng.uniform(0.2, 0.8, (a, 2))
    wh = rng.uniform(0.05, 0.3, (a, 2))
    boxes = np.concatenate([centres, wh], 1).astype(np.float32)
    scores = rng.permutation(a).astype(np.float32) / a * 0.9 + 0.05          # distinct, > 0.01
    probs = np.zeros((1, a, 2), np.float32)
    probs[0, :, 0] = scores
    probs[0, :, 1] = 1 - scores
    kept, count = M.nms(probs, boxes[None], score_thresh=0.01, top_k=200, iou_thresh=0.45)
    order = np.argsort(-scores, kind="stable")[:200]
    xyxy = np.concatenate([boxes[:, :2] - boxes[:, 2:] / 2, boxes[:, :2] + boxes[:, 2:] / 2], 1)
    iou = tv.ops.box_iou(torch.from_numpy(xyxy[order]), torch.from_numpy(xyxy[order])).numpy()
    near = np.abs(iou - 0.45) < 1e-4
    assert not near.any()                                                   # the two IoU formulas cannot disagree here
    tv_keep = tv.ops.nms(torch.from_numpy(xyxy[order]), torch.from_numpy(scores[order]), 0.45).numpy()
    assert np.array_equal(kept[0, 0, :count[0, 0]], order[tv_keep])


# ---- autograd bridge (reference: tape.gradient through _ssd_loss, models/ssd_model.py:240-248) --------------------
def test_autograd_bridge_gradients_flow_without_host_round_trips(priors300):
    torch = pytest.importorskip("torch")
    from ssdgeom import autograd as AG
    b = 3
    boxes, cls, off = synth.make_gt(51, b, 40, "coco")
    y = _targets(boxes, cls, off, priors300, b)
    pc, pb = synth.make_predictions(51, b, 8732)
    dev = torch.device("cuda")
    t_cls = torch.from_numpy(pc).to(dev).requires_grad_(True)
    t_box = torch.from_numpy(pb).to(dev).requires_grad_(True)
    y_t = (torch.from_numpy(y[0]).to(dev), torch.from_numpy(y[1]).to(dev), torch.from_numpy(y[2]).to(dev))
    total, info = AG.ssd_loss(y_t, (t_box, t_cls))
    assert total.is_cuda and total.dtype == torch.float64 and total.requires_grad
    (total * 2.0).backward()                  # upstream gradient 2: the bridge scales by it
    w_total, w_info = O.ssd_loss((y[0], y[1], y[2].astype(bool)), (pb, pc))
    np.testing.assert_allclose(total.item(), w_total, rtol=1e-5)
    for k in ("cls loss pos", "cls loss neg", "loc loss"):
        np.testing.assert_allclose(info[k].item(), w_info[k], rtol=1e-5)
    o_box, o_cls = O.ssd_loss_grad((y[0], y[1], y[2].astype(bool)), (pb, pc))
    np.testing.assert_allclose(t_cls.grad.cpu().numpy(), 2.0 * o_cls, rtol=2e-5, atol=1e-9)
    np.testing.assert_allclose(t_box.grad.cpu().numpy(), 2.0 * o_box, rtol=1e-5, atol=1e-12)
    # a gradient step on the logits lowers the loss (sanity of the sign)
    with torch.no_grad():
        t_cls2 = (t_cls - 50.0 * t_cls.grad / 2.0).detach().requires_grad_(False)
    total2, _ = AG.ssd_loss(y_t, (t_box.detach(), t_cls2))
    assert total2.item() < total.item()


# ---- get_train_set's prefetch (models/ssd_model.py:225) -----------------------------------------------------------
def test_train_batches_prefetch_matches_synchronous_iteration(priors300):
    from ssdgeom import data
    rng = np.random.default_rng(2)
    boxes, cls, off = synth.make_gt(61, 10, 15, "coco")

    def source():
        for i in range(10):
            yield (rng.uniform(size=(8, 8, 3)).astype(np.float32), cls[off[i]:off[i + 1]], boxes[off[i]:off[i + 1]])

    rng = np.random.default_rng(2)
    sync_batches = list(data.TrainBatches(source(), priors300, batch_size=4, prefetch=0))
    rng = np.random.default_rng(2)
    pre_batches = list(data.TrainBatches(source(), priors300, batch_size=4, prefetch=3))
    assert len(sync_batches) == len(pre_batches) == 2                       # drop_remainder
    for (ia, ta), (ib, tb) in zip(sync_batches, pre_batches):
        assert np.array_equal(ia, ib)
        for x, y in zip(ta, tb):
            assert np.array_equal(x, y)
    w = O.assign_encode(cls[off[0]:off[1]], boxes[off[0]:off[1]], priors300, 0.5, sweeps=False)
    assert np.array_equal(pre_batches[0][1][0][0], w[0]) and np.array_equal(pre_batches[0][1][2][0], w[2])


# ---- candidate lists in prior space: the run layout must not show in the results -----------------------------
@pytest.mark.parametrize("bias", [7.0, 0.0])
def test_detect_is_independent_of_the_batch_it_runs_in(bias, priors300):
    """The filter CTAs own contiguous runs of tiles and leave per-run counts; how an image is cut into runs depends
    on the batch (1 image: 137 runs of 2 tiles; 5 images: runs that straddle images; bias 0 makes lists longer than
    the sort width, i.e. the radix select reads its list through the runs several times).  Same images, same kept
    lists and scores, bit for bit."""
    pred_cls, pred_box = synth.make_predictions(977, 5, 8732, bg_bias=bias)
    for aux in (False, True):   # the streaming variant of the filter pass, and the one that also writes the side outputs
        full = M.detect(pred_cls, pred_box, priors300, return_aux=aux)
        for lo, hi in ((0, 1), (1, 2), (2, 3), (3, 4), (4, 5), (1, 4)):
            part = M.detect(pred_cls[lo:hi], pred_box[lo:hi], priors300, return_aux=aux)
            assert np.array_equal(part[1], full[1][lo:hi]) and np.array_equal(part[0], full[0][lo:hi])
            if aux:
                assert np.array_equal(part[2]["boxes"], full[2]["boxes"][lo:hi])
                assert np.array_equal(part[2]["kept_score"], full[2]["kept_score"][lo:hi])


# ---- two steps in flight: alternate buffers, no result may depend on it ------------------------------------------
def test_hotpath_two_steps_in_flight_equal_closed_steps(priors300):
    """HotPath(depth=2) lets step k+1 start under the tail of step k on a second set of per-step buffers.  Seven steps
    over three different resident batches, results fetched only where a consumer would fetch them (after every
    step, and once after a burst of steps): loss block, kept lists, counts and targets equal the closed steps'."""
    b, gtmax = 6, 30
    gts = [synth.make_gt(300 + k, b, gtmax, "max") for k in range(3)]
    total = max(g[0].shape[0] for g in gts)
    preds = [synth.make_predictions(50 + 10 * k, b, 8732) for k in range(3)]

    def build(depth):
        hp = HotPath(synth.SSD300, batch=b, max_gt=gtmax, total_gt=total, depth=depth)
        for k in range(3):
            if k:
                hp.add_input_set()
            hp.use_set(k)
            boxes, cls, off = gts[k]
            pad = total - boxes.shape[0]
            hp.upload(np.concatenate([boxes, np.zeros((pad, 4), np.float32)]), np.concatenate([cls, np.zeros((pad,), np.float32)]),
                      off, *preds[k])
        hp.s_main.sync()
        return hp

    def fetch(hp):
        out = (np.zeros(N.LOSS_RESULT_LEN), np.zeros((b, 80, 200), np.int32), np.zeros((b, 80), np.int32))
        hp.download(*out)
        hp.s_main.sync()
        return out + (hp.tgt["cls"].to_host(), hp.tgt["mask"].to_host(), hp.tgt["loc"].to_host())

    order = [0, 1, 2, 1, 0, 2, 2]
    closed, hp1 = [], build(1)
    for k in order:
        hp1.use_set(k)
        hp1.step()
        closed.append(fetch(hp1))
    hp2 = build(2)
    for n, k in enumerate(order):        # fetched after every step
        hp2.use_set(k)
        hp2.step()
        for x, y in zip(fetch(hp2), closed[n]):
            assert np.array_equal(x, y)
    for k in order:                      # a burst: nothing waits in between
        hp2.use_set(k)
        hp2.step()
    for x, y in zip(fetch(hp2), closed[-1]):
        assert np.array_equal(x, y)
    hp2.check_status()


def test_detect_many_tiny_images_run_geometry():
    """4000 images of 24 priors: one tile per image, so a filter CTA's run holds dozens of images -- the shared-memory
    counter budget caps the run length (more CTAs than SMs) and every image is one short run.  The batch must give
    what its slices give, and the oracle's answer on a few images."""
    rng = np.random.default_rng(5)
    b, a, c = 4000, 24, 81
    pri = np.concatenate([rng.uniform(0.1, 0.9, (a, 2)), rng.uniform(0.05, 0.4, (a, 2))], 1)
    pred_cls = rng.normal(size=(b, a, c)).astype(np.float32)
    pred_cls[..., -1] += 2.0
    pred_box = (rng.normal(size=(b, a, 4)) * 0.3).astype(np.float32)
    kw = dict(score_thresh=0.01, top_k=32, iou_thresh=0.45)
    kept, count = M.detect(pred_cls, pred_box, pri, **kw)
    for lo, hi in ((0, 40), (1777, 1800), (3990, 4000)):
        k2, c2 = M.detect(pred_cls[lo:hi], pred_box[lo:hi], pri, **kw)
        assert np.array_equal(c2, count[lo:hi]) and np.array_equal(k2, kept[lo:hi])
    for i in (0, 2345, 3999):
        w_kept, w_count, _, _ = O.detect(pred_cls[i], pred_box[i], pri, **kw)
        same = sum(int(np.array_equal(kept[i, cc, :count[i, cc]], w_kept[cc, :w_count[cc]])) for cc in range(c - 1))
        assert same >= c - 1 - 2      # near-ties of the float32 scores aside (see the A9 tests)


# ---- BASELINE config 4's table on two shards ------------------------------------------------------------------
def test_ssd512_two_shards_chained_step_against_the_oracle():
    """SSD512 (24 564 priors), a batch of 4 images cut into 2 shards by image as the data-parallel path cuts it
    (parallel.shard_csr): every shard runs the chained step (HotPath) on its slice.  Targets per image equal the
    oracle's bit for bit, each shard's loss is the oracle's _ssd_loss of that slice (per-shard mining: the reference's
    split_batch semantics, models/ssd_model.py:235-256), the pooled combination of the additive sums equals the one
    computed from the oracle's per-shard terms, and the detections equal the single-device call on the whole batch."""
    t512 = synth.TABLES["ssd512"]
    priors = O.build_prior_box(t512["sizes"], t512["s_k_refer"], t512["aspect_ratio"], t512["input_size"])
    a = priors.shape[0]
    assert a == 24564
    batch, world = 4, 2
    boxes, cls, off = synth.make_gt(512, batch, 60, "max")
    pred_cls, pred_box = synth.make_predictions(513, batch, a, bg_bias=7.0)
    kept_all, count_all = M.detect(pred_cls, pred_box, priors)
    sums, want_sums = [], []
    for rank in range(world):
        (lo, hi), (r0, r1), soff = parallel.shard_csr(off, world, rank)
        n = hi - lo
        hp = HotPath(t512, batch=n, max_gt=60, total_gt=r1 - r0)
        out = (np.zeros(N.LOSS_RESULT_LEN), np.zeros((n, 80, 200), np.int32), np.zeros((n, 80), np.int32))
        hp.step_host(boxes[r0:r1], cls[r0:r1], soff, pred_cls[lo:hi], pred_box[lo:hi], *out)
        g_cls, g_loc, g_mask = hp.tgt["cls"].to_host(), hp.tgt["loc"].to_host(), hp.tgt["mask"].to_host()
        w = [O.assign_encode(cls[off[i]:off[i + 1]], boxes[off[i]:off[i + 1]], priors, 0.5, sweeps=False) for i in range(lo, hi)]
        for k in range(n):
            assert np.array_equal(g_cls[k], w[k][0]) and np.array_equal(g_mask[k], w[k][2])
            np.testing.assert_allclose(g_loc[k], w[k][1], rtol=1e-5, atol=1e-7)
        y_true = (np.stack([x[0] for x in w]), np.stack([x[1] for x in w]).astype(np.float32), np.stack([x[2] for x in w]))
        w_total, w_info, w_aux = O.ssd_loss(y_true, (pred_box[lo:hi], pred_cls[lo:hi]), return_masks=True)
        np.testing.assert_allclose(out[0][0], w_total, rtol=1e-5)
        assert out[0][4] == w_aux["num_pos"] and out[0][5] == w_aux["num_neg"]
        # the chained step's detections are the streaming filter variant of the stand-alone call
        assert np.array_equal(out[2], count_all[lo:hi]) and np.array_equal(out[1], kept_all[lo:hi])
        sums.append(parallel.block_sums(out[0]))
        want_sums.append({"num_pos": w_aux["num_pos"], "num_neg": w_aux["num_neg"],
                          "sum_pos_ce": w_info["cls loss pos"] * w_aux["num_pos"],
                          "sum_neg_ce": w_info["cls loss neg"] * w_aux["num_neg"],
                          "sum_l1": w_info["loc loss"] * w_aux["num_pos"]})
    got, _ = parallel.combine_loss(sums, "pooled")
    want, _ = parallel.combine_loss(want_sums, "pooled")
    np.testing.assert_allclose(got, want, rtol=1e-5)
