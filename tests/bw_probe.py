#!/usr/bin/env python
"""Bandwidth reference points on the box (write-only, read-only, copy) next to the decode kernel."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gq_b200  # noqa: E402
from gq_b200 import _lib  # noqa: E402
from util import codebook  # noqa: E402

dev = torch.device("cuda", 0)
N = 23498432
bufs = [torch.empty(N, device=dev) for _ in range(4)]
src = [torch.randn(N, device=dev) for _ in range(4)]


def timeit(fn, iters=40):
    for i in range(5):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


ms = timeit(lambda i: bufs[i % 4].fill_(1.0))
print("fill   94 MB: %.4f ms  %.0f GB/s (write only)" % (ms, N * 4 / ms / 1e6))
ms = timeit(lambda i: bufs[i % 4].copy_(src[(i + 1) % 4]))
print("copy   94 MB: %.4f ms  %.0f GB/s (read+write)" % (ms, N * 8 / ms / 1e6))
ms = timeit(lambda i: src[i % 4].sum())
print("sum    94 MB: %.4f ms  %.0f GB/s (read only)" % (ms, N * 4 / ms / 1e6))

n_chunks = N // 16
cb = torch.from_numpy(codebook(16, 256)).to(dev)
codes = torch.randint(0, 256, (n_chunks,), dtype=torch.uint8, device=dev)
l = torch.randint(0, 65, (n_chunks,), dtype=torch.uint8, device=dev)
lbub = torch.tensor([-0.05, 0.05], device=dev)
seg = torch.tensor([0, n_chunks], dtype=torch.int64, device=dev)
for U, mean in ((1, 0), (1, 1)):
    ms = timeit(lambda i: _lib.call("gq_hsq_decode_reduce", codes.data_ptr(), 1, l.data_ptr(), 1, lbub.data_ptr(),
                                    None, 0, U, n_chunks, 16, cb.data_ptr(), 256, seg.data_ptr(), 1, 6, mean, 0,
                                    bufs[i % 4].data_ptr(), _lib.stream()))
    print("decode U=%d mean=%d: %.4f ms  %.0f GB/s algorithmic" % (U, mean, ms, (N * 4 + n_chunks * 2) / ms / 1e6))
# U users: records laid out with a common stride
for U in (2, 4, 8):
    stride = ((2 * n_chunks + 8 + 255) // 256) * 256
    rec = torch.randint(0, 65, (U, stride), dtype=torch.uint8, device=dev)
    rec[:, 2 * n_chunks:2 * n_chunks + 8] = torch.tensor([-0.05, 0.05], device=dev).view(torch.uint8).repeat(U, 1)
    base = rec.data_ptr()
    ms = timeit(lambda i: _lib.call("gq_hsq_decode_reduce", base, 1, base + n_chunks, 1, base + 2 * n_chunks, None, stride,
                                    U, n_chunks, 16, cb.data_ptr(), 256, seg.data_ptr(), 1, 6, 1, 0,
                                    bufs[i % 4].data_ptr(), _lib.stream()))
    print("decode U=%d mean=1: %.4f ms  %.0f GB/s algorithmic" % (U, ms, (N * 4 + U * n_chunks * 2) / ms / 1e6))
u = torch.randn(n_chunks, device=dev)
lq = torch.empty(n_chunks, dtype=torch.uint8, device=dev)
keys = torch.empty(2, dtype=torch.int32, device=dev)
lb2 = torch.empty(2, device=dev)
ms = timeit(lambda i: _lib.call("gq_norm_quantize", u.data_ptr(), n_chunks, seg.data_ptr(), 1, 6, 1, None, 1, 4 * i,
                                lq.data_ptr(), 1, lb2.data_ptr(), keys.data_ptr(), 0, _lib.stream()))
print("norm_quantize (minmax + quantize, philox) %d chunks: %.4f ms" % (n_chunks, ms))
