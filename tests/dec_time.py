#!/usr/bin/env python
"""Timing helper (GPU box): decode-and-average of U users' packed ResNet-50 records held locally.
GQ_DECODE_STAGED=1 forces the pull-and-decode (shared-memory staged) kernel."""
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
for U in ([int(x) for x in sys.argv[1:]] or [1, 2, 4, 8]):
    a = make_args(mode="ps", num_users=U)
    params = [torch.nn.Parameter(torch.zeros(s, device=dev)) for s in shapes]
    q = gq_b200.Quantizer(gq_b200.NearestNeighborCompressor, params, a)
    plan = q.plan
    for u in range(U):
        plan.arena.normal_(0, 0.01)
        plan.encode(u)
    outs = [torch.empty_like(plan.arena) for _ in range(3)]
    for i in range(5):
        plan.decode(n_users=U, mean=True, out=outs[i % 3])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(60):
        plan.decode(n_users=U, mean=True, out=outs[i % 3])
    e1.record()
    torch.cuda.synchronize()
    print("staged=%s U=%d decode %.4f ms" % (os.environ.get("GQ_DECODE_STAGED", "-"), U, e0.elapsed_time(e1) / 60),
          flush=True)
    del q, plan, params, outs
