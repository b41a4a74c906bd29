import sys, numpy as np, ctypes as C
sys.path.insert(0,'ssd-object-detection_b200'); sys.path.insert(0,'.')
from ssdgeom import synth, ops, device as D
from ssdgeom.models import ssd_model as M
pri = ops.prior_boxes(M.SSD300["sizes"], M.SSD300["s_k_refer"], M.SSD300["aspect_ratio"], 300)
ops.prior_index(pri)
b,c,o = synth.make_gt(100, 256, 100, "max")
for it in range(3):
    out = ops.match_encode(b,c,o,pri,256,100,0.5)
    D.sync()
ws = out["_match_ws"]
head = ws.view((64,), np.uint32).to_host()
t = head[8:8+10].view(np.uint64)
print("phase clocks per image (setup, search, phase2, greedy, output):", [int(x)//256 for x in t])
w = head[24:32].view(np.uint64)
rows = 256*100
print("per row: pass1 clk %d, evaluations clk %d (n=%.2f), pass2 tests clk %d" % (w[0]//rows, w[1]//rows, w[3]/rows, w[2]//rows))

print("re-search events per image %.2f, rows per image %.2f, clocks per image %d" % (head[40]/256, head[41]/256, int(head[42:44].view(np.uint64)[0])//256))
