#!/usr/bin/env python
"""Diagnostics for the tcgen05 kernel: which of (v read from smem, approximate row max) is
wrong for mismatching chunks, and where (CTA, local iteration, row)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gq_b200  # noqa: E402,F401
from gq_b200 import _lib  # noqa: E402
from oracle import gq_oracle as O  # noqa: E402
from util import codebook, gen_input  # noqa: E402

dev = torch.device("cuda", 0)
cb = codebook(16, 256)
cbt = torch.from_numpy(cb).to(dev)
n_chunks = int(sys.argv[1]) if len(sys.argv) > 1 else 56909
x = gen_input(123, n_chunks * 16, "normal")
xt = torch.from_numpy(x).to(dev)
codes = torch.full((n_chunks,), 255, dtype=torch.uint8, device=dev)
u = torch.zeros(n_chunks, device=dev)
seg = torch.tensor([0, n_chunks], dtype=torch.int64, device=dev)
dbg = torch.zeros(128 * 256 + n_chunks * 24, device=dev)
_lib.call("gq_hsq_tc_debug", xt.data_ptr(), n_chunks, cbt.data_ptr(), codes.data_ptr(), u.data_ptr(),
          seg.data_ptr(), 1, dbg.data_ptr(), 1, _lib.stream())
torch.cuda.synchronize()
oc, ou = O.hsq_search(x.reshape(-1, 16), cb)
gc = codes.cpu().numpy().astype(np.int32)
gu = u.cpu().numpy()
aux = dbg.cpu().numpy()[128 * 256:].reshape(-1, 24)
bad = (gc != oc) | (gu != ou)
v0_wrong = aux[:, 0] != x.reshape(-1, 16)[:, 0]
amax_rel = np.abs(aux[:, 1] - np.abs(ou)) / np.maximum(np.abs(ou), 1e-30)
print("n_chunks %d grid %s: bad %d, v0 wrong %d, amax off by >1%% %d" %
      (n_chunks, os.environ.get("GQ_TC_GRID", "default"), bad.sum(), v0_wrong.sum(), (amax_rel > 0.01).sum()))
print("bad & v0 wrong: %d   bad & amax off: %d" % ((bad & v0_wrong).sum(), (bad & (amax_rel > 0.01)).sum()))
xs = x.reshape(-1, 16)
vs_true = np.zeros(n_chunks, np.float32)
for j in range(16):
    vs_true = vs_true + xs[:, j]
print("flags=%s  vsum wrong: %d (of bad: %d)" % (os.environ.get("GQ_TC_FLAGS", "0"), (aux[:, 4] != vs_true).sum(), ((aux[:, 4] != vs_true) & bad).sum()))
cs_true = np.zeros(256, np.float32)
cbs = cb.copy()
# the kernel sums the physical row (a permutation of 16-byte units): float add is order dependent,
# so compare loosely
cs_true = cbs.astype(np.float64).sum(1)
print("codeword row-sum off: %d" % (np.abs(aux[:, 5] - cs_true[gc]) > 1e-5).sum())
masks = aux[:, 6].view(np.uint32)
og = oc // 4
in_mask = np.where(og < 32, ((masks >> (og % 32).astype(np.uint32)) & 1).astype(bool), True)
print("oracle winner group missing from candidate mask: %d (of bad %d)   mean popcount %.3f" % ((~in_mask).sum(), (~in_mask & bad).sum(), np.mean([bin(m).count("1") for m in masks[:20000]])))
print("bad with same group as oracle: %d / %d" % ((bad & ((gc // 4) == og)).sum(), bad.sum()))
if bad.any():
    its = aux[bad, 2].astype(int)
    print("local iteration histogram of bad chunks (it: count):", dict(zip(*np.unique(its, return_counts=True))))
    rows = np.flatnonzero(bad) % 128
    print("row-quad histogram:", np.bincount(rows // 32, minlength=4))
    ctas = aux[bad, 3].astype(int)
    print("distinct CTAs with bad chunks:", len(np.unique(ctas)))
    # does the wrong v0 equal the v0 of the chunk 6 local iterations later (stage overwritten)?
    idx = np.flatnonzero(bad & v0_wrong)[:2000]
    grid = int(aux[:, 3].max()) + 1
    later = idx + 6 * grid * 128
    ok = later < n_chunks
    same = (aux[idx[ok], 0] == x.reshape(-1, 16)[later[ok], 0]).mean() if ok.any() else float("nan")
    print("fraction of wrong v0 equal to the same row 6 iterations later: %.3f" % same)

vd = aux[:, 8:24]
wrong = vd != xs
print("per-index wrong counts (j=0..15):", wrong.sum(0))
bi = np.flatnonzero(bad)[:6]
ex = O.hsq_scores(xs[bi], cb)
for n, i in enumerate(bi):
    print("chunk", i, "it", int(aux[i, 2]), "row", i % 128)
    print("   v kernel:", vd[i])
    print("   v true  :", xs[i])
    w = np.flatnonzero(wrong[i])
    # are the wrong values scores of this row (TMEM data)?
    for j in w[:4]:
        d = np.abs(ex[n] - vd[i, j])
        print("     v[%d]=%.6e nearest exact score of this row: k=%d diff %.2e ; equals x elsewhere? %s" % (j, vd[i, j], d.argmin(), d.min(), np.flatnonzero(x == vd[i, j])[:3]))
