#!/usr/bin/env python
"""A/B of the decode-and-average kernels at U users (GPU box): staged (four lanes per chunk) vs owner
(one lane per chunk), results compared bit for bit, then timed.   python tests/dec_ab.py [U ...]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gq_b200  # noqa: E402
from util import make_args, resnet50_shapes  # noqa: E402

dev = torch.device("cuda", 0)
shapes = resnet50_shapes()
ok = True
for U in ([int(x) for x in sys.argv[1:]] or [3, 4, 8]):
    a = make_args(mode="ps", num_users=U)
    params = [torch.nn.Parameter(torch.zeros(s, device=dev)) for s in shapes]
    q = gq_b200.Quantizer(gq_b200.NearestNeighborCompressor, params, a)
    plan = q.plan
    for u in range(U):
        plan.arena.normal_(0, 0.01)
        plan.encode(u)
    outs = [torch.empty_like(plan.arena) for _ in range(3)]
    res = {}
    for mode in ("0", "1"):
        os.environ["GQ_DECODE_OWNER"] = mode
        o = torch.zeros_like(plan.arena)
        plan.decode(n_users=U, mean=True, out=o)
        plan.decode(n_users=U, mean=False, accumulate=1, out=o)     # out += sum
        torch.cuda.synchronize()
        res[mode] = o
        for i in range(5):
            plan.decode(n_users=U, mean=True, out=outs[i % 3])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(60):
            plan.decode(n_users=U, mean=True, out=outs[i % 3])
        e1.record()
        torch.cuda.synchronize()
        print("U=%d owner=%s decode %.1f us" % (U, mode, e0.elapsed_time(e1) / 60 * 1e3), flush=True)
    same = torch.equal(res["0"].view(torch.int32), res["1"].view(torch.int32))
    ok = ok and same
    print("U=%d owner vs staged: %s" % (U, "identical" if same else "MISMATCH"), flush=True)
    del q, plan, params, outs, res
print("DECAB %s" % ("OK" if ok else "FAILED"))
sys.exit(0 if ok else 1)
