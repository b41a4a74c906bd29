"""CPU: the multi-rank host logic (image sharding + the single all-reduce of the additive loss sums)
under torch.distributed with the gloo backend, world_size 2.  The per-shard numbers come from the
oracle here (no GPU); on the GPU box the same functions consume the kernels' result blocks."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ssd_oracle as O
from ssdgeom import parallel, synth


def test_shard_ranges_cover_and_balance():
    for n in (1, 7, 256, 1000):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_range(4, 2, 2)


def test_shard_csr():
    boxes, cls, off = synth.make_gt(3, 7, 20, "coco")
    seen = 0
    for r in range(3):
        (lo, hi), (r0, r1), sub = parallel.shard_csr(off, 3, r)
        assert sub[0] == 0 and sub[-1] == r1 - r0 and np.array_equal(np.diff(sub), np.diff(off[lo:hi + 1]))
        seen += hi - lo
    assert seen == 7


def _shard_block(y_true, y_pred, lo, hi):
    yt = tuple(v[lo:hi] for v in y_true)
    yp = tuple(v[lo:hi] for v in y_pred)
    total, info, aux = O.ssd_loss(yt, yp, return_masks=True)
    block = np.zeros(16)
    block[0:4] = [total, info["cls loss pos"], info["cls loss neg"], info["loc loss"]]
    block[4], block[5] = aux["num_pos"], aux["num_neg"]
    block[8] = info["cls loss pos"] * aux["num_pos"]
    block[9] = info["cls loss neg"] * aux["num_neg"]
    block[10] = info["loc loss"] * aux["num_pos"]
    return block


def _problem():
    rng = np.random.default_rng(4)
    b, a, c = 6, 300, 9
    gt_cls = rng.integers(0, c - 1, (b, a)).astype(np.int32)
    gt_mask = rng.uniform(size=(b, a)) < 0.05
    gt_mask[:, 0] = True
    gt_box = rng.normal(size=(b, a, 4)).astype(np.float32)
    pred_box = rng.normal(size=(b, a, 4)).astype(np.float32)
    pred_cls = rng.normal(size=(b, a, c)).astype(np.float32)
    return (gt_cls, gt_box, gt_mask), (pred_box, pred_cls)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        y_true, y_pred = _problem()
        lo, hi = parallel.shard_range(y_true[0].shape[0], world, rank)
        block = torch.from_numpy(_shard_block(y_true, y_pred, lo, hi))
        pooled = parallel.distributed_loss(block, mode="pooled")
        mean = parallel.distributed_loss(block, mode="mean_of_shards")
        out[rank] = (pooled, mean, block.numpy().copy())
    finally:
        dist.destroy_process_group()


def test_two_rank_loss_exchange_gloo():
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    y_true, y_pred = _problem()
    blocks = [_shard_block(y_true, y_pred, *parallel.shard_range(6, world, r)) for r in range(world)]
    want_pooled = parallel.combine_loss([parallel.block_sums(b) for b in blocks], "pooled")
    want_mean = parallel.combine_loss([parallel.block_sums(b) for b in blocks], "mean_of_shards")
    for r in range(world):
        pooled, mean, block = res[r]
        assert np.array_equal(block, blocks[r])                 # the exchange does not clobber the local block
        np.testing.assert_allclose(pooled[0], want_pooled[0], rtol=1e-12)
        np.testing.assert_allclose(mean[0], want_mean[0], rtol=1e-12)
        for k in ("cls loss pos", "cls loss neg", "loc loss"):
            np.testing.assert_allclose(pooled[1][k], want_pooled[1][k], rtol=1e-12)
            np.testing.assert_allclose(mean[1][k], want_mean[1][k], rtol=1e-12)
    # both ranks agree, and the reference's own split-batch semantics (mean of per-slice losses) is reproduced
    assert res[0][0][0] == res[1][0][0] and res[0][1][0] == res[1][1][0]
    per_slice = [O.ssd_loss(tuple(v[lo:hi] for v in y_true), tuple(v[lo:hi] for v in y_pred))[0]
                 for lo, hi in (parallel.shard_range(6, world, r) for r in range(world))]
    np.testing.assert_allclose(want_mean[0], np.mean(per_slice), rtol=1e-12)


# ---- exact batch-global mining across shards --------------------------------------------------------------
class _HostStagedLoss:
    """NumPy stand-in for ops.StagedLoss (same run / exchange / finish protocol and the same exchange words:
    three 2048-bin histograms over the order-preserving key, 11 + 11 + 10 bits, and the positives count), fed
    by the oracle's per-prior CE.  It lets the protocol driver run under gloo without a GPU."""

    def __init__(self, y_true, y_pred, global_priors, ratio=3):
        gt_cls, gt_box, gt_mask = y_true
        pred_box, pred_cls = y_pred
        self.pos = gt_mask.astype(bool)
        bg = np.full(gt_cls.shape, pred_cls.shape[-1] - 1, dtype=np.int64)
        self.neg_ce = (O.softmax_ce(pred_cls, bg) * (~self.pos).astype(np.float32)).reshape(-1)
        self.ce_gt = O.softmax_ce(pred_cls, gt_cls)
        self.l1 = np.abs(pred_box.astype(np.float32) - gt_box.astype(np.float32)).astype(np.float64).sum(-1)
        bits = self.neg_ce.view(np.uint32).astype(np.uint64)
        self.key = np.where(bits >> 31 != 0, ~bits & 0xffffffff, bits | 0x80000000).astype(np.uint64)
        self.hist = [np.zeros(2048, np.int32) for _ in range(3)]
        self.npos = np.zeros(1, np.int64)
        self.sums = np.zeros(4)
        self.nneg = np.zeros(1)
        self.ratio, self.global_priors = ratio, global_priors

    def _state(self, levels):
        k = self.ratio * int(self.npos[0])
        assert 1 <= k <= self.global_priors
        prefix = 0
        for lv, (shift, nb) in enumerate(((21, 2048), (10, 2048), (0, 1024))[:levels]):
            above = 0
            for b in range(nb - 1, -1, -1):
                if above + int(self.hist[lv][b]) >= k:
                    break
                above += int(self.hist[lv][b])
            k -= above
            prefix |= b << shift
        return prefix

    def run(self, stage):
        if stage == 0:
            self.npos[0] = int(self.pos.sum())
            self.hist[0][:] = np.bincount((self.key >> 21).astype(np.int64), minlength=2048)
        elif stage == 1:
            sel = (self.key >> 21) == (self._state(1) >> 21)
            self.hist[1][:] = np.bincount(((self.key[sel] >> 10) & 2047).astype(np.int64), minlength=2048)
        elif stage == 2:
            sel = (self.key >> 10) == (self._state(2) >> 10)
            self.hist[2][:] = np.bincount((self.key[sel] & 1023).astype(np.int64), minlength=2048)
        else:
            self.kth_key = self._state(3)
            self.neg_mask = self.key >= self.kth_key
            self.sums[:] = [self.ce_gt[self.pos].sum(dtype=np.float64), self.neg_ce[self.neg_mask].sum(dtype=np.float64),
                            self.l1[self.pos].sum(), self.pos.sum()]
            self.nneg[0] = self.neg_mask.sum()

    def exchange(self, stage):
        return [[self.npos, self.hist[0]], [self.hist[1]], [self.hist[2]], [self.sums, self.nneg]][stage]

    def finish(self):
        s_pos, s_neg, s_l1, n_pos = self.sums
        return parallel.loss_from_sums(s_pos, s_neg, s_l1, n_pos, self.nneg[0])


def _global_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        y_true, y_pred = _problem()
        b, a = y_true[0].shape
        lo, hi = parallel.shard_range(b, world, rank)
        staged = _HostStagedLoss(tuple(v[lo:hi] for v in y_true), tuple(v[lo:hi] for v in y_pred), b * a)
        total, info = parallel.global_mining_loss(staged, parallel.torch_allreduce())
        out[rank] = (total, info, staged.neg_mask.reshape(hi - lo, a).copy())
    finally:
        dist.destroy_process_group()


def test_two_rank_exact_global_mining_gloo():
    """The staged protocol reproduces the single-shard loss and negative mask of the WHOLE batch
    (models/ssd_model.py:368-372 with the batch spread over two ranks)."""
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_global_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    y_true, y_pred = _problem()
    want_total, want_info, aux = O.ssd_loss(y_true, y_pred, return_masks=True)
    masks = np.concatenate([res[r][2] for r in range(world)])
    assert np.array_equal(masks, aux["neg_mask"])
    for r in range(world):
        np.testing.assert_allclose(res[r][0], want_total, rtol=1e-12)
        for k in ("cls loss pos", "cls loss neg", "loc loss"):
            np.testing.assert_allclose(res[r][1][k], want_info[k], rtol=1e-12)
    assert res[0][0] == res[1][0]
    # per-shard mining (the default) differs on this problem: the test is not vacuous
    per = [_shard_block(y_true, y_pred, *parallel.shard_range(6, world, r)) for r in range(world)]
    assert parallel.combine_loss([parallel.block_sums(b) for b in per], "pooled")[0] != want_total
