#!/usr/bin/env python
"""Pipeline trace of the second-generation tcgen05 search kernel (CTA 0, first 128 tiles).

    python tests/tc2_trace.py [variant]      (GQ_TC2 switch string, default pair,fmask,f2)

Prints, per epilogue group and averaged over the steady-state tiles, how many cycles a tile spends in
each phase -- the evidence behind DESIGN.md's statement of what bounds the kernel."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
if len(sys.argv) > 1:
    os.environ["GQ_TC2"] = sys.argv[1]
os.environ["GQ_TC_V"] = "2"
import gq_b200  # noqa: E402,F401
from gq_b200 import _lib  # noqa: E402
from util import codebook  # noqa: E402

dev = torch.device("cuda", 0)
cbt = torch.from_numpy(codebook(16, 256)).to(dev)
n = 1468652
xs = [torch.randn(n * 16, device=dev) * 0.01 for _ in range(3)]
codes = torch.empty(n, dtype=torch.uint8, device=dev)
u = torch.empty(n, device=dev)
GRID = 148
trace = torch.zeros(16 * 128 + 4 * GRID, dtype=torch.int64, device=dev)
full = "--encode" in sys.argv
lv = torch.empty(n, dtype=torch.uint8, device=dev)
# a ResNet-50-like segment table: 76 tensors of equal size
n_seg = 76
seg = torch.tensor([n * i // n_seg // 4 * 4 for i in range(n_seg)] + [n], dtype=torch.int64, device=dev)
lbub = torch.empty(2 * n_seg, device=dev)
ws = torch.zeros(1 << 16, dtype=torch.uint8, device=dev)
for i in range(3):
    _lib.call("gq_hsq_tc2_trace", xs[i].data_ptr(), n, cbt.data_ptr(), codes.data_ptr(), u.data_ptr(), seg.data_ptr(), n_seg,
              lv.data_ptr() if full else None, lbub.data_ptr(), ws.data_ptr(), trace.data_ptr(), _lib.stream())
torch.cuda.synchronize()
raw = trace.cpu().numpy().astype(np.int64)
t = raw[:16 * 128].reshape(16, 128)
st = raw[16 * 128:].reshape(GRID, 4)
st = st[st[:, 0] > 0]
t_first = st[:, 0].min()
print("per-CTA wall clock (us since the first CTA started), %d CTAs%s:" % (st.shape[0], " [whole encode]" if full else " [search only]"))
for k, nm in enumerate(["start", "main loop done", "grid barrier passed", "tail done"]):
    col = (st[:, k] - t_first) / 1e3
    if (st[:, k] > 0).all():
        print("  %-20s min %7.2f  median %7.2f  max %7.2f" % (nm, col.min(), np.median(col), col.max()))
tiles = 77
t0 = t[0, 0]
ev = (t[:8, :tiles] - t0)
names = ["tma", "mma", "acc seen", "tmem rel", "stage rel", "mask done", "rescored", "done"]
print("first 12 tiles (cycles since the first TMA):")
for it in range(12):
    print("  tile %2d: " % it + "  ".join("%s %6d" % (nm, ev[k, it]) for k, nm in enumerate(names)) + "  iters %d" % t[8, it])
lo, hi = 9, tiles - 3
sl = slice(lo, hi)
print("steady state (tiles %d..%d), mean cycles:" % (lo, hi - 1))
print("  tile period (done[i+3] - done[i]) / 3 : %.0f" % np.mean((ev[7, lo + 3:hi + 3] - ev[7, lo:hi]) / 3.0))
print("  wait for accumulators (acc seen - previous tile of the group done): %.0f" % np.mean(ev[2, lo + 3:hi + 3] - ev[7, lo:hi]))
print("  first pass  (tmem rel - acc seen) : %.0f" % np.mean(ev[3, sl] - ev[2, sl]))
print("  v + norm    (stage rel - tmem rel): %.0f" % np.mean(ev[4, sl] - ev[3, sl]))
print("  max + mask  (mask done - stage rel): %.0f" % np.mean(ev[5, sl] - ev[4, sl]))
print("  rescoring   (rescored - mask done): %.0f   iterations %.2f" % (np.mean(ev[6, sl] - ev[5, sl]), np.mean(t[8, sl])))
print("  store/minmax (done - rescored)    : %.0f" % np.mean(ev[7, sl] - ev[6, sl]))
print("  MMA issue -> accumulators seen    : %.0f" % np.mean(ev[2, sl] - ev[1, sl]))
print("  TMEM release(i) -> MMA issue(i+2) : %.0f" % np.mean(ev[1, lo + 2:hi + 2] - ev[3, lo:hi]))
print("  TMA issue -> MMA issue            : %.0f" % np.mean(ev[1, sl] - ev[0, sl]))
print("  total for the CTA: %d cycles for %d tiles" % (ev[7, tiles - 1], tiles))

# round-2 diagnostics: where the TMEM buffer turnaround goes (all four quadrants, MMA warp, polling observer)
rel_all = np.stack([t[3, :tiles], t[9, :tiles], t[10, :tiles], t[11, :tiles]]) - t0
last_rel = rel_all.max(axis=0)
print("  quad release skew (last quad - quad 0)        : %.0f   (last quad - first quad: %.0f)" % (np.mean(last_rel[sl] - ev[3, sl]), np.mean(last_rel[sl] - rel_all.min(axis=0)[sl])))
first_rel = rel_all.min(axis=0)
print("  per-quadrant release - first release (mean)   : " + "  ".join("q%d %.0f" % (q, np.mean(rel_all[q, sl] - first_rel[sl])) for q in range(4)))
print("  which quadrant is last (share of tiles)       : " + "  ".join("q%d %.0f%%" % (q, 100.0 * np.mean(rel_all[:, sl].argmax(axis=0) == q)) for q in range(4)))
for grp in range(3):
    idx = np.arange(lo + ((grp - lo) % 3), hi, 3)
    print("  group %d: quadrant release - first (mean)       : " % grp + "  ".join("q%d %.0f" % (q, np.mean(rel_all[q, idx] - first_rel[idx])) for q in range(4)))
m_fu = t[12, :tiles] - t0
m_te = t[13, :tiles] - t0
obs = t[14, :tiles] - t0
print("  last release(i) -> MMA warp past tempty(i+2)  : %.0f" % np.mean(m_te[lo + 2:hi + 2] - last_rel[lo:hi]))
print("  MMA warp: tempty passed -> issued             : %.0f" % np.mean(ev[1, sl] - m_te[sl]))
print("  MMA issue -> commit observed by polling warp  : %.0f" % np.mean(obs[sl] - ev[1, sl]))
print("  commit observed -> epilogue quad 0 sees it    : %.0f" % np.mean(ev[2, sl] - obs[sl]))
print("  MMA warp idle (prev issue -> tempty passed)   : %.0f" % np.mean(m_te[lo + 1:hi + 1] - ev[1, lo:hi]))
