#!/usr/bin/env python
"""Stand-alone probe of the tcgen05 search kernel (run on the GPU box under `timeout`):
correctness vs the oracle, the measured TF32 approximation error vs the rescoring margin,
and a first timing.  Not a pytest file (a protocol bug would trap the process)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gq_b200  # noqa: E402
from gq_b200 import _lib  # noqa: E402
from oracle import gq_oracle as O  # noqa: E402
from util import codebook, gen_input  # noqa: E402

dev = torch.device("cuda", 0)
cb = codebook(16, 256)
cbt = torch.from_numpy(cb).to(dev)


def run(n_chunks, kind, dbg_tiles=0):
    x = gen_input(123, n_chunks * 16, kind)
    xt = torch.from_numpy(x).to(dev)
    codes = torch.full((n_chunks,), 255, dtype=torch.uint8, device=dev)
    u = torch.zeros(n_chunks, device=dev)
    seg = torch.tensor([0, n_chunks], dtype=torch.int64, device=dev)
    dbg = torch.zeros(max(dbg_tiles, 1) * 128 * 256, device=dev)
    _lib.call("gq_hsq_tc_debug", xt.data_ptr(), n_chunks, cbt.data_ptr(), codes.data_ptr(), u.data_ptr(),
              seg.data_ptr(), 1, dbg.data_ptr() if dbg_tiles else None, dbg_tiles, _lib.stream())
    torch.cuda.synchronize()
    oc, ou = O.hsq_search(x.reshape(-1, 16), cb)
    gc = codes.cpu().numpy().astype(np.int32)
    gu = u.cpu().numpy()
    bad = int((gc != oc).sum())
    badu = int((gu != ou).sum())
    print("n_chunks=%d kind=%s: code mismatches %d, u mismatches %d" % (n_chunks, kind, bad, badu), flush=True)
    if bad:
        idx = np.flatnonzero(gc != oc)[:10]
        print("   first bad chunks", idx, "gpu", gc[idx], "oracle", oc[idx], flush=True)
    if dbg_tiles:
        rows = min(dbg_tiles * 128, n_chunks)
        approx = dbg.cpu().numpy().reshape(-1, 256)[:rows]
        exact = O.hsq_scores(x.reshape(-1, 16)[:rows], cb)
        nv = np.linalg.norm(x.reshape(-1, 16)[:rows].astype(np.float64), axis=1)
        err = np.abs(approx.astype(np.float64) - exact.astype(np.float64)).max(1) / np.maximum(nv, 1e-300)
        eps = 1.5 / 1024 + 4e-6   # hsq_tc.cu kMargin / 2 (unit-norm codebook)
        print("   TF32 score error / ||v||: max %.3e  mean %.3e   (eps used: %.3e, margin 2eps %.3e)"
              % (err.max(), err.mean(), eps, 2 * eps), flush=True)
        assert err.max() < eps, "approximation error exceeds the bound the rescoring margin is built on"
        print("   approx-argmax == exact-argmax on %.4f of rows"
              % (np.abs(approx).argmax(1) == np.abs(exact).argmax(1)).mean(), flush=True)
    return bad + badu


fails = 0
fails += run(128, "normal", dbg_tiles=1)
fails += run(128 * 7 + 5, "normal", dbg_tiles=8)
fails += run(128 * 148 * 3 + 77, "heavy", dbg_tiles=16)
fails += run(1468652, "normal")

# timing: ResNet-50-sized search, exact vs tensor core
n_chunks = 1468652
xs = [torch.randn(n_chunks * 16, device=dev) * 0.01 for _ in range(4)]
codes = torch.empty(n_chunks, dtype=torch.uint8, device=dev)
u = torch.empty(n_chunks, device=dev)
seg = torch.tensor([0, n_chunks], dtype=torch.int64, device=dev)
ws = torch.empty(1 << 20, dtype=torch.uint8, device=dev)
for name, algo in (("exact", _lib.ALGO_EXACT), ("tc", _lib.ALGO_TC)):
    for i in range(3):
        _lib.call("gq_hsq_search", xs[i % 4].data_ptr(), n_chunks, 16, cbt.data_ptr(), 256, codes.data_ptr(), 1,
                  u.data_ptr(), seg.data_ptr(), 1, None, ws.data_ptr(), ws.numel(), algo, _lib.stream())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        _lib.call("gq_hsq_search", xs[i % 4].data_ptr(), n_chunks, 16, cbt.data_ptr(), 256, codes.data_ptr(), 1,
                  u.data_ptr(), seg.data_ptr(), 1, None, ws.data_ptr(), ws.numel(), algo, _lib.stream())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print("search %-5s: %.3f ms  -> %.1f Gelem/s, %.0f GB/s algorithmic" %
          (name, ms, n_chunks * 16 / ms / 1e6, n_chunks * (64 + 5) / ms / 1e6), flush=True)
print("PROBE", "FAILED" if fails else "OK", flush=True)
sys.exit(1 if fails else 0)
