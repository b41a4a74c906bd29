"""Multi-GPU check of the exact batch-global mining, one process per GPU over NCCL:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/check_global_mining.py

Every rank builds the same full batch, computes the single-device loss of the WHOLE batch, then runs the staged
loss on its own slice with the exchange words all-reduced over NCCL, and asserts that its mined mask equals its
slice of the single-device mask bit for bit, that threshold and counts are identical, and that the total agrees
to 1e-9 (the per-lane float partial sums depend on how the priors fall into tiles).  Prints one line per rank; exit code 0 on success."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "ssd-object-detection_b200"))
from ssdgeom import _native as N, ops, parallel, synth          # noqa: E402
from ssdgeom.models import ssd_model as M                        # noqa: E402


def main():
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    torch.cuda.set_device(local)
    N.check(N.lib().ssdg_set_device(local), "set_device")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    batch = 8 * world
    priors = M.build_prior_box(M.SSD300["sizes"])
    boxes, cls, off = synth.make_gt(77, batch, 100, "coco")
    tgt = ops.match_encode(boxes, cls, off, priors, batch, 100, 0.5)
    y = [tgt[k].to_host() for k in ("cls", "loc", "mask")]
    pred_cls, pred_box = synth.make_predictions(77, batch, priors.shape[0])
    full = ops.multibox_loss(y[0], y[1], y[2], pred_box, pred_cls, want_neg_mask=True)
    want = ops.loss_result_to_host(full["result"])
    want_mask = full["neg_mask"].to_host()
    lo, hi = parallel.shard_range(batch, world, rank)
    staged = ops.StagedLoss(y[0][lo:hi], y[1][lo:hi], y[2][lo:hi], pred_box[lo:hi], pred_cls[lo:hi],
                            global_priors=batch * priors.shape[0], want_neg_mask=True)
    total, info = parallel.global_mining_loss(staged, parallel.torch_allreduce())
    ok = (np.array_equal(staged.out["neg_mask"].to_host(), want_mask[lo:hi]) and info["kth"] == want["kth"]
          and info["num_neg"] == want["num_neg"] and info["num_pos"] == want["num_pos"]
          and abs(total - want["total"]) <= 1e-9 * abs(want["total"]))
    print("rank %d/%d mask_equal %s images [%d,%d) total %.12f want %.12f num_neg %d kth %.9g -> %s" %
          (rank, world, np.array_equal(staged.out["neg_mask"].to_host(), want_mask[lo:hi]), lo, hi, total, want["total"], info["num_neg"], info["kth"], "OK" if ok else "MISMATCH"), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
