#!/usr/bin/env python
"""tcgen05 HSQ encode at chunk dimensions 8 and 32 (K = 256): bit-exactness against the exact CUDA-core
kernel on edge-case inputs and on a 25.6 M-element gradient (codes, u, and the whole packed record of the
one-launch encode), the C oracle on a sample, then timings.  GPU box only.

    python tests/tc2_dims.py [8] [32] [16]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import gq_b200  # noqa: E402
from gq_b200 import _lib  # noqa: E402
from gq_b200.quantizers.fused import FusedPlan  # noqa: E402
from oracle import gq_oracle as O  # noqa: E402
from util import codebook, gen_input, make_args  # noqa: E402

dev = torch.device("cuda", 0)
dims = [int(a) for a in sys.argv[1:] if a.isdigit()] or [8, 32]
ws = torch.empty(1 << 20, dtype=torch.uint8, device=dev)
fails = 0


def search(x, d, cbt, algo):
    n = x.numel() // d
    codes = torch.full((n,), 255, dtype=torch.uint8, device=dev)
    u = torch.full((n,), 7.0, device=dev)
    seg = torch.tensor([0, n], dtype=torch.int64, device=dev)
    _lib.call("gq_hsq_search", x.data_ptr(), n, d, cbt.data_ptr(), 256, codes.data_ptr(), 1, u.data_ptr(),
              seg.data_ptr(), 1, None, ws.data_ptr(), ws.numel(), algo, _lib.stream())
    torch.cuda.synchronize()
    return codes, u


def check(name, x, d, cbt):
    global fails
    c1, u1 = search(x, d, cbt, _lib.ALGO_EXACT)
    c2, u2 = search(x, d, cbt, _lib.ALGO_TC)
    bad_c = int((c1 != c2).sum())
    bad_u = int((u1.view(torch.int32) != u2.view(torch.int32)).sum())
    if bad_c or bad_u:
        fails += 1
        idx = torch.nonzero(c1 != c2).flatten()[:8].tolist()
        print("  MISMATCH d=%d %s: codes %d u-bits %d first %s" % (d, name, bad_c, bad_u, idx), flush=True)
    else:
        print("  ok d=%d %s (%d chunks)" % (d, name, x.numel() // d), flush=True)


def time_loop(fn, iters=20):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


for d in dims:
    cb = codebook(d, 256)
    cbt = torch.from_numpy(cb).to(dev)
    for kind in ("normal", "heavy", "zeros_mixed"):
        for n_chunks in (1, 127, 128, 129, 128 * 7 + 5, 128 * 148 * 3 + 77):
            x = torch.from_numpy(gen_input(11 + n_chunks % 97, n_chunks * d, kind)).to(dev)
            check("%s/%d" % (kind, n_chunks), x, d, cbt)
    x = torch.from_numpy(gen_input(5, 128 * 40 * d, "normal")).to(dev).view(-1, d)
    x[0:700] *= 1e-20
    x[700:1400] *= 1e-34
    x[1400:2100] *= 1e30
    x[2100:2110] = 0.0
    x[2110:2120] = -0.0
    x[2120, 3] = float("inf")
    x[2121, 5] = float("nan")
    x[2122] = 3.0e38
    x[2123:2200] *= 1e-12
    check("scales", x.reshape(-1).contiguous(), d, cbt)
    cw = cbt.repeat(20, 1) * torch.linspace(-2, 2, 20 * 256, device=dev).view(-1, 1)
    check("codewords", cw.reshape(-1).contiguous(), d, cbt)
    # oracle on a sample
    xs = gen_input(77, 4096 * d, "normal")
    c2, u2 = search(torch.from_numpy(xs).to(dev), d, cbt, _lib.ALGO_TC)
    oc, ou = O.hsq_search(xs.reshape(-1, d), cb)
    if not (np.array_equal(c2.cpu().numpy().astype(np.int64), np.asarray(oc).astype(np.int64)) and
            np.array_equal(u2.cpu().numpy().view(np.int32), np.asarray(ou, dtype=np.float32).view(np.int32))):
        fails += 1
        print("  MISMATCH d=%d vs oracle" % d, flush=True)
    else:
        print("  ok d=%d oracle sample" % d, flush=True)

    # full-size flat gradient through the plan: packed record of the one-launch encode vs the exact path
    shapes = [(25_600_000,)]
    plans = {}
    for name, algo in (("tc", _lib.ALGO_AUTO), ("exact", _lib.ALGO_EXACT)):
        plans[name] = FusedPlan(gq_b200.NearestNeighborCompressor, shapes, make_args(c_dim=d, num_users=1, hsq_algo=algo), dev, 1)
    p = plans["tc"]
    gen = torch.Generator(device=dev)
    gen.manual_seed(7)
    g = torch.randn(p.arena_elems, device=dev, generator=gen) * 0.01
    n = p.groups[0].n_chunks
    uni = {id(plans[k].groups[0]): torch.rand(n, device=dev, generator=torch.Generator(device=dev).manual_seed(3)) for k in plans}
    for k in plans:
        plans[k].records.zero_()
        plans[k].encode(0, src=g, uniforms=uni)
    torch.cuda.synchronize()
    same = torch.equal(plans["tc"].records[0], plans["exact"].records[0])
    same_u = torch.equal(plans["tc"].u_scratch[:n].view(torch.int32), plans["exact"].u_scratch[:n].view(torch.int32))
    if not (same and same_u):
        fails += 1
        grp = p.groups[0]
        a, b = plans["tc"].records[0], plans["exact"].records[0]
        for nm, off, ln in (("codes", grp.codes_off, n), ("l", grp.l_off, n), ("lbub", grp.lbub_off, 8 * grp.n_seg)):
            print("  MISMATCH d=%d full-size %s: %d bytes differ" % (d, nm, int((a[off:off + ln] != b[off:off + ln]).sum())), flush=True)
        print("  u equal: %s" % same_u, flush=True)
    else:
        print("  ok d=%d full-size record (25.6 M elements, %d chunks)" % (d, n), flush=True)
    # timing over rotating inputs (> L2)
    inputs = [torch.randn(p.arena_elems, device=dev) * 0.01 for _ in range(3)]
    out = torch.empty(p.arena_elems, device=dev)
    t_tc = time_loop(lambda i: plans["tc"].encode(0, src=inputs[i % 3]))
    t_ex = time_loop(lambda i: plans["exact"].encode(0, src=inputs[i % 3]), iters=5)
    t_dec = time_loop(lambda i: plans["tc"].decode(first_user=0, n_users=1, mean=True, out=out))
    gb = p.groups[0].n * 4 / 1e3
    print("DIM d=%d ok=%d encode tcgen05 %.1f us (%.0f GB/s in) exact %.1f us decode %.1f us" %
          (d, 0 if fails else 1, t_tc, gb / t_tc, t_ex, t_dec), flush=True)
    del plans, p, inputs, out, g, uni
    torch.cuda.empty_cache()
print("TC2DIMS %s" % ("OK" if not fails else "FAILED"), flush=True)
sys.exit(1 if fails else 0)
