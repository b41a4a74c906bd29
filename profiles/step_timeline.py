"""Timeline of one chained step: begin / end of every bracketed kernel relative to the step's first event
(CUDA events on the launching streams).  usage: python profiles/step_timeline.py [batch]"""
import ctypes as C, os, sys, statistics
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "ssd-object-detection_b200"))
from ssdgeom import _native as N, device as D, synth   # noqa: E402
from ssdgeom.pipeline import HotPath                   # noqa: E402

b = int(sys.argv[1]) if len(sys.argv) > 1 else 256
boxes, cls, off = synth.make_gt(100, b, 100, "max")
hp = HotPath(synth.TABLES["ssd300"], batch=b, max_gt=100, total_gt=boxes.shape[0])
pc = np.empty((b, hp.A, hp.classes), np.float32); pb = np.empty((b, hp.A, 4), np.float32)
for i in range(0, b, 16):
    n = min(16, b - i)
    pc[i:i + n], pb[i:i + n] = synth.make_predictions(i, n, hp.A, hp.classes)
hp.upload(boxes, cls, off, pc, pb); hp.s_main.sync()
for _ in range(3):
    hp.step()
hp.s_main.sync()
N.lib().ssdg_profile_enable(1)
names = {N.PROF_FILTER: "filter", N.PROF_NMS: "nms", N.PROF_SEARCH: "search",
         N.PROF_MATCH: "search+match", N.PROF_CE: "ce/lossprep", N.PROF_LOSS_TAIL: "select+final"}
spans = {k: [] for k in names}
tot = []
e1 = D.Event()
for _ in range(12):
    hp.step(); e1.record(hp.s_main); hp.s_main.sync()
    tot.append(hp.ev_begin.elapsed_ms(e1) * 1e3)
    for k in names:
        a, z = C.c_float(0), C.c_float(0)
        if N.lib().ssdg_profile_span_ms(k, hp.ev_begin.handle, C.byref(a), C.byref(z)) == 0:
            spans[k].append((a.value * 1e3, z.value * 1e3))
N.lib().ssdg_profile_enable(0)
print("step %.1f us" % statistics.median(tot[2:]))
for k, v in sorted(spans.items(), key=lambda kv: statistics.median([x[0] for x in kv[1][2:]]) if kv[1] else 0):
    if v:
        print("%-14s %8.1f -> %8.1f us" % (names[k], statistics.median([x[0] for x in v[2:]]), statistics.median([x[1] for x in v[2:]])))
