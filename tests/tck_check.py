#!/usr/bin/env python
"""Stand-alone check + timing of the large-codebook tcgen05 search (hsq_tck.cu), run on the GPU box
under `timeout` (a protocol bug traps the context): codes and u against the exact CUDA-core kernel,
bit for bit, for K = 512 / 1024 / 2048 (synthesised unit-norm codebooks) and the learned K = 4096
codebook, on edge-case sizes and inputs; then the ResNet-50-size timing of both kernels."""
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
fails = 0


def search(x, cbt, K, algo):
    n = x.numel() // 16
    codes = torch.full((n,), -1, dtype=torch.int32, device=dev)
    u = torch.full((n,), 7.0, device=dev)
    seg = torch.tensor([0, n], dtype=torch.int64, device=dev)
    nb = _lib.value("gq_hsq_encode_workspace_bytes", n, 16, K, 1)
    ws = torch.empty(nb + 8192, dtype=torch.uint8, device=dev)
    _lib.call("gq_hsq_search", x.data_ptr(), n, 16, cbt.data_ptr(), K, codes.data_ptr(), 4, u.data_ptr(), seg.data_ptr(), 1,
              None, ws.data_ptr(), ws.numel(), algo, _lib.stream())
    torch.cuda.synchronize()
    return codes, u


def check(name, x, cbt, K):
    global fails
    c1, u1 = search(x, cbt, K, _lib.ALGO_EXACT)
    c2, u2 = search(x, cbt, K, _lib.ALGO_TC)
    bad_c = int((c1 != c2).sum())
    bad_u = int((u1.view(torch.int32) != u2.view(torch.int32)).sum())
    if bad_c or bad_u:
        fails += 1
        idx = torch.nonzero(c1 != c2).flatten()[:6].tolist()
        print("  MISMATCH %s: codes %d u-bits %d first %s tc %s exact %s" %
              (name, bad_c, bad_u, idx, c2[idx].tolist(), c1[idx].tolist()), flush=True)
    else:
        print("  ok %s (%d chunks)" % (name, x.numel() // 16), flush=True)


books = {}
for K in (512, 1024, 2048):
    books[K] = torch.from_numpy(O.normalize(np.random.RandomState(K).standard_normal((K, 16)).astype(np.float32))[1]).to(dev)
books[4096] = torch.from_numpy(codebook(16, 4096)).to(dev)
for K, cbt in books.items():
    for kind in ("normal", "heavy", "zeros_mixed"):
        for n_chunks in (1, 127, 129, 128 * 7 + 5, 128 * 148 * 2 + 77):
            x = torch.from_numpy(gen_input(11 + n_chunks % 97, n_chunks * 16, kind)).to(dev)
            check("K=%d %s/%d" % (K, kind, n_chunks), x, cbt, K)
    x = torch.from_numpy(gen_input(5, 128 * 20 * 16, "normal")).to(dev).view(-1, 16)
    x[0:300] *= 1e-20
    x[300:600] *= 1e30
    x[600:610] = 0.0
    x[610:620] = -0.0
    x[620, 3] = float("inf")
    x[621, 5] = float("nan")
    x[622:700] *= 1e-12
    check("K=%d scales" % K, x.reshape(-1).contiguous(), cbt, K)
    cw = cbt[:512].repeat(4, 1) * torch.linspace(-2, 2, 2048, device=dev).view(-1, 1)
    check("K=%d codewords" % K, cw.reshape(-1).contiguous(), cbt, K)

# oracle spot check (K = 4096, learned codebook)
x = gen_input(3, 16 * 5000)
c2, u2 = search(torch.from_numpy(x).to(dev), books[4096], 4096, _lib.ALGO_TC)
oc, ou = O.hsq_search(x.reshape(-1, 16), codebook(16, 4096))
ok = np.array_equal(c2.cpu().numpy(), oc) and np.array_equal(u2.cpu().numpy(), ou)
print("  %s oracle K=4096" % ("ok" if ok else "MISMATCH"), flush=True)
fails += 0 if ok else 1

# timing at ResNet-50 size
n = 1468652
xs = [torch.randn(n * 16, device=dev) * 0.01 for _ in range(3)]
codes = torch.empty(n, dtype=torch.int32, device=dev)
u = torch.empty(n, device=dev)
seg = torch.tensor([0, n], dtype=torch.int64, device=dev)
for K in (4096, 1024):
    nb = _lib.value("gq_hsq_encode_workspace_bytes", n, 16, K, 1)
    ws = torch.empty(nb + 8192, dtype=torch.uint8, device=dev)
    for name, algo, iters in (("tcgen05", _lib.ALGO_TC, 10), ("exact", _lib.ALGO_EXACT, 3)):
        def run(i):
            _lib.call("gq_hsq_search", xs[i % 3].data_ptr(), n, 16, books[K].data_ptr(), K, codes.data_ptr(), 4, u.data_ptr(),
                      seg.data_ptr(), 1, None, ws.data_ptr(), ws.numel(), algo, _lib.stream())
        run(0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            run(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        print("TCK K=%d %-8s %.3f ms  %.1f Gelem/s  %.1f TFLOP/s" % (K, name, ms, n * 16 / ms / 1e6, 2.0 * K * n * 16 / ms / 1e9),
              flush=True)
print("TCK", "FAILED" if fails else "OK", flush=True)
sys.exit(1 if fails else 0)
