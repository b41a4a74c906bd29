"""Which pairs of stages actually overlap?  Times each stage alone and the pairs the pipeline co-schedules
(CUDA events on s_main, medians over `reps` runs).  usage: python profiles/pipeline_overlap.py [batch]"""
import os, sys, statistics
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "ssd-object-detection_b200"))
from ssdgeom import device as D, synth          # noqa: E402
from ssdgeom.pipeline import HotPath            # noqa: E402

b = int(sys.argv[1]) if len(sys.argv) > 1 else 256
boxes, cls, off = synth.make_gt(100, b, 100, "max")
hp = HotPath(synth.TABLES["ssd300"], batch=b, max_gt=100, total_gt=boxes.shape[0])
pc = np.empty((b, hp.A, hp.classes), np.float32); pb = np.empty((b, hp.A, 4), np.float32)
for i in range(0, b, 16):
    n = min(16, b - i)
    pc[i:i + n], pb[i:i + n] = synth.make_predictions(i, n, hp.A, hp.classes)
hp.upload(boxes, cls, off, pc, pb); hp.s_main.sync()
hp.step(); hp.s_main.sync()

def timed(label, branches, reps=15):
    """branches: list of (stream, [callables(stream)]) started together after ev_begin."""
    ts = []
    e0, e1 = D.Event(), D.Event()
    evs = [D.Event() for _ in branches]
    for _ in range(reps):
        e0.record(hp.s_main)
        for st, _ in branches:
            D.stream_wait_event(st, e0)
        for (st, fns), ev in zip(branches, evs):
            for fn in fns:
                fn(st)
            ev.record(st)
        for ev in evs:
            D.stream_wait_event(hp.s_main, ev)
        e1.record(hp.s_main)
        hp.s_main.sync()
        ts.append(e0.elapsed_ms(e1) * 1e3)
    print("%-34s %8.1f us" % (label, statistics.median(ts[3:])), flush=True)

filt = lambda s: hp.detect_stage(s, stage=0, stats=True)
nms = lambda s: hp.detect_stage(s, stage=1, stats=True)
loss = lambda s: hp.loss_stage(s, stats=True)
timed("filter", [(hp.s_d, [filt])])
timed("match", [(hp.s_a, [hp.assign])])
timed("filter || match", [(hp.s_d, [filt]), (hp.s_a, [hp.assign])])
timed("nms", [(hp.s_n, [nms])])
timed("loss", [(hp.s_l, [loss])])
timed("nms(low) || loss(high)", [(hp.s_n, [nms]), (hp.s_l, [loss])])
timed("nms(high) || loss(low)", [(hp.s_d, [nms]), (hp.s_a, [loss])])
timed("nms || match", [(hp.s_n, [nms]), (hp.s_a, [hp.assign])])
timed("filter || loss", [(hp.s_d, [filt]), (hp.s_l, [loss])])
timed("filter+nms || match+loss", [(hp.s_d, [filt, nms]), (hp.s_a, [hp.assign, loss])])

def full():
    e0, e1 = D.Event(), D.Event()
    ts = []
    for _ in range(15):
        e0.record(hp.s_main); hp.step(); e1.record(hp.s_main); hp.s_main.sync(); ts.append(e0.elapsed_ms(e1) * 1e3)
    return statistics.median(ts[3:])
print("%-34s %8.1f us" % ("HotPath.step (split, fused)", full()))
