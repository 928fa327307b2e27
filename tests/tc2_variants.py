#!/usr/bin/env python
"""A/B harness for the tcgen05 HSQ encode kernels (run on the GPU box, one variant per process:
a protocol bug traps the context).

    python tests/tc2_variants.py <variant> [--quick]
        variant: "v1" (hsq_tc.cu + quantize launch) or a GQ_TC2 switch string, e.g. "g3,pair,fmask,f2"

Checks, against the exact CUDA-core kernel on the same inputs: codes and u of the search, and the
whole packed record (codes, l, lb/ub, identity) of the one-launch encode with external uniforms --
bit for bit, on edge-case inputs and on the full ResNet-50 gradient.  Then times search-only and
encode with CUDA events over rotating inputs (> L2).  Prints one line: VARIANT <name> ok=<0|1> ...
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

variant = sys.argv[1]
quick = "--quick" in sys.argv
if variant == "v1":
    os.environ["GQ_TC_V"] = "1"
else:
    os.environ["GQ_TC_V"] = "2"
    os.environ["GQ_TC2"] = variant

import gq_b200  # noqa: E402
from gq_b200 import _lib  # noqa: E402
from gq_b200.quantizers.fused import FusedPlan  # noqa: E402
from util import codebook, gen_input, make_args, resnet50_shapes  # noqa: E402

dev = torch.device("cuda", 0)
cbt = torch.from_numpy(codebook(16, 256)).to(dev)
ws = torch.empty(1 << 20, dtype=torch.uint8, device=dev)
fails = 0


def search(x, algo, n_seg_tab=None):
    n = x.numel() // 16
    codes = torch.full((n,), 255, dtype=torch.uint8, device=dev)
    u = torch.full((n,), 7.0, device=dev)
    seg = torch.tensor([0, n], dtype=torch.int64, device=dev) if n_seg_tab is None else n_seg_tab
    _lib.call("gq_hsq_search", x.data_ptr(), n, 16, cbt.data_ptr(), 256, codes.data_ptr(), 1, u.data_ptr(),
              seg.data_ptr(), seg.numel() - 1, None, ws.data_ptr(), ws.numel(), algo, _lib.stream())
    torch.cuda.synchronize()
    return codes, u


def check_search(name, x):
    global fails
    c1, u1 = search(x, _lib.ALGO_EXACT)
    c2, u2 = search(x, _lib.ALGO_TC)
    bad_c = int((c1 != c2).sum())
    bad_u = int((u1.view(torch.int32) != u2.view(torch.int32)).sum())
    if bad_c or bad_u:
        fails += 1
        idx = torch.nonzero(c1 != c2).flatten()[:8].tolist()
        print("  MISMATCH %s: codes %d u-bits %d first %s" % (name, bad_c, bad_u, idx), flush=True)
    else:
        print("  ok %s (%d chunks)" % (name, x.numel() // 16), flush=True)


# ---- edge-case inputs ----
for kind in ("normal", "heavy", "zeros_mixed", "repeat16"):
    for n_chunks in (1, 127, 128, 129, 128 * 7 + 5, 128 * 148 * 3 + 77):
        x = torch.from_numpy(gen_input(11 + n_chunks % 97, n_chunks * 16, kind)).to(dev)
        check_search("%s/%d" % (kind, n_chunks), x)
# scales: denormal range, huge, mixed signs of zero, inf / nan rows
x = torch.from_numpy(gen_input(5, 128 * 40 * 16, "normal")).to(dev).view(-1, 16)
x[0:700] *= 1e-20
x[700:1400] *= 1e-34
x[1400:2100] *= 1e30
x[2100:2110] = 0.0
x[2110:2120] = -0.0
x[2120, 3] = float("inf")
x[2121, 5] = float("nan")
x[2122] = 3.0e38
x[2123:2200] *= 1e-12
check_search("scales", x.reshape(-1).contiguous())
# codeword inputs (exact ties between +c and -c never happen; near-ties between groups do)
cw = cbt.repeat(20, 1) * torch.linspace(-2, 2, 20 * 256, device=dev).view(-1, 1)
check_search("codewords", cw.reshape(-1).contiguous())

# ---- full-size plan: record bytes of the one-launch encode vs the exact path ----
shapes = resnet50_shapes()
plans = {}
for name, algo in (("tc", _lib.ALGO_AUTO), ("exact", _lib.ALGO_EXACT)):
    plans[name] = FusedPlan(gq_b200.NearestNeighborCompressor, shapes, make_args(num_users=1, hsq_algo=algo), dev, 1)
p = plans["tc"]
gen = torch.Generator(device=dev)
gen.manual_seed(7)
g = torch.randn(p.arena_elems, device=dev, generator=gen) * 0.01
for i in range(len(p.sizes)):
    v = p.view(i, g)
    v.mul_(float(10.0 ** ((i % 7) - 4)))
    if i % 37 == 5:
        v.zero_()
n = p.groups[0].n_chunks
for random in (1, 0):
    uni = {id(plans[k].groups[0]): torch.rand(n, device=dev, generator=torch.Generator(device=dev).manual_seed(3))
           for k in plans}
    for k in plans:
        plans[k].random = random
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
            d = int((a[off:off + ln] != b[off:off + ln]).sum())
            print("  MISMATCH full-size random=%d %s: %d bytes differ" % (random, nm, d), flush=True)
        print("  u equal: %s" % same_u, flush=True)
    else:
        print("  ok full-size record random=%d" % random, flush=True)
# Philox path (no external uniforms): tc vs exact must agree as well (same stream, same indexing)
for k in plans:
    plans[k].random = 1
    torch.manual_seed(1234)
    plans[k].records.zero_()
    plans[k].encode(0, src=g)
torch.cuda.synchronize()
if not torch.equal(plans["tc"].records[0], plans["exact"].records[0]):
    fails += 1
    print("  MISMATCH full-size Philox record", flush=True)
else:
    print("  ok full-size Philox record", flush=True)

# ---- timing ----
ROT = 4
inputs = [torch.randn(p.arena_elems, device=dev) * 0.01 for _ in range(ROT)]
grp = p.groups[0]
st = _lib.stream()


def time_loop(fn, iters=40):
    for i in range(5):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def search_only(i):
    _lib.call("gq_hsq_search", inputs[i % ROT].data_ptr() + grp.arena_off * 4, grp.n_chunks, 16, grp.codebook.data_ptr(),
              256, p.records[0].data_ptr() + grp.codes_off, 1, p.u_scratch.data_ptr(), grp.seg_start.data_ptr(),
              grp.n_seg, None, p.workspace.data_ptr() + 4096, p.workspace.numel() - 4096, _lib.ALGO_TC, st)


out = torch.empty(p.arena_elems, device=dev)
t_search = time_loop(search_only)
t_encode = time_loop(lambda i: p.encode(0, src=inputs[i % ROT]))
t_step = time_loop(lambda i: (p.encode(0, src=inputs[i % ROT]), p.decode(first_user=0, n_users=1, mean=True, out=out)))
t_decode = time_loop(lambda i: p.decode(first_user=0, n_users=1, mean=True, out=out))
print("VARIANT %-22s ok=%d search_us=%.1f encode_us=%.1f decode_us=%.1f step_us=%.1f" %
      (variant, 0 if fails else 1, t_search, t_encode, t_decode, t_step), flush=True)
sys.exit(1 if fails else 0)
