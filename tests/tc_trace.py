#!/usr/bin/env python
"""Pipeline timeline of CTA 0 of the tcgen05 search kernel (debug build time-stamps)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gq_b200
from gq_b200 import _lib
from util import codebook
dev = torch.device("cuda", 0)
cbt = torch.from_numpy(codebook(16, 256)).to(dev)
n_chunks = 1468652
x = torch.randn(n_chunks * 16, device=dev) * 0.01
codes = torch.empty(n_chunks, dtype=torch.uint8, device=dev); u = torch.empty(n_chunks, device=dev)
seg = torch.tensor([0, n_chunks], dtype=torch.int64, device=dev)
dbg = torch.zeros(n_chunks * 24 + 2 * 6 * 128 + 16, device=dev)
for _ in range(2):
    _lib.call("gq_hsq_tc_debug", x.data_ptr(), n_chunks, cbt.data_ptr(), codes.data_ptr(), u.data_ptr(),
              seg.data_ptr(), 1, dbg.data_ptr(), 0, _lib.stream())
torch.cuda.synchronize()
tr = dbg[n_chunks * 24: n_chunks * 24 + 2 * 6 * 128].cpu().numpy().view(np.int64).reshape(6, 128)
t0 = tr[0, 0]
names = ["TMA issued", "MMA issued", "acc seen", "TMEM released", "stage released", "tile done"]
print("it  " + "  ".join("%14s" % n for n in names) + "   (cycles since first TMA issue; debug build)")
for it in list(range(0, 24)) + list(range(60, 70)):
    print("%3d " % it + "  ".join("%14d" % (tr[e, it] - t0) for e in range(6)))
d = np.diff(tr[1, 10:70]); print("MMA issue period: mean %.0f clk" % d.mean())
d = np.diff(tr[5, 10:70]); print("tile done period (per group every 3rd): mean %.0f clk" % d.mean())
print("acc seen - MMA issued: mean %.0f" % (tr[2, 10:70] - tr[1, 10:70]).mean())
print("TMEM released - acc seen: mean %.0f" % (tr[3, 10:70] - tr[2, 10:70]).mean())
print("stage released - TMEM released: mean %.0f" % (tr[4, 10:70] - tr[3, 10:70]).mean())
print("tile done - stage released: mean %.0f" % (tr[5, 10:70] - tr[4, 10:70]).mean())
print("MMA issued(it+2) - TMEM released(it): mean %.0f" % (tr[1, 12:72] - tr[3, 10:70]).mean())
print("MMA issued(it) - TMA issued(it): mean %.0f" % (tr[1, 10:70] - tr[0, 10:70]).mean())
