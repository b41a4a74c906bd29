import os, sys, time, statistics
import numpy as np
sys.path.insert(0, os.path.join(os.environ.get("GRAFT_REPO_ROOT", "."), "ssd-object-detection_b200"))
from ssdgeom import device as D, synth
from ssdgeom.pipeline import HotPath
b = int(sys.argv[1]) if len(sys.argv) > 1 else 256
boxes, cls, off = synth.make_gt(100, b, 100, "max")
hp = HotPath(synth.TABLES["ssd300"], batch=b, max_gt=100, total_gt=boxes.shape[0])
pc = np.empty((b, hp.A, hp.classes), np.float32); pb = np.empty((b, hp.A, 4), np.float32)
for i in range(0, b, 16):
    n = min(16, b - i)
    pc[i:i + n], pb[i:i + n] = synth.make_predictions(i, n, hp.A, hp.classes)
hp.upload(boxes, cls, off, pc, pb); hp.s_main.sync()
for _ in range(5): hp.step()
hp.s_main.sync()
ts = []
for _ in range(30):
    hp.s_main.sync()
    t0 = time.perf_counter(); hp.step(); t1 = time.perf_counter()
    ts.append((t1 - t0) * 1e6)
print("batch", b, "host enqueue per step: median %.1f us  min %.1f  max %.1f" % (statistics.median(ts), min(ts), max(ts)))
