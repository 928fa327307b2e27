#!/usr/bin/env python
"""Timing helper (torchrun, N >= 2 GPUs): cost of the pieces of the peer-to-peer ps exchange on the
ResNet-50 record -- barrier kernel, gather kernel, local decode, pull-and-decode."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gq_b200  # noqa: E402
from util import make_args, resnet50_shapes  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
a = make_args(mode="ps", num_users=world)
params = [torch.nn.Parameter(torch.zeros(s, device=dev)) for s in resnet50_shapes()]
q = gq_b200.Quantizer(gq_b200.NearestNeighborCompressor, params, a)
assert q.p2p is not None
plan, p2p = q.plan, q.p2p
out = torch.empty_like(plan.arena)
if rank == 0:
    print("alloc=%s multicast=%s" % (p2p.alloc, bool(p2p.mc_base)), flush=True)
for row in range(2 * world):          # fill every row of the local block
    plan.arena.normal_(0, 0.01)
    plan.encode(row)
torch.cuda.synchronize()
dist.barrier()


def timed(name, fn, iters=60):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("%-34s %.4f ms" % (name, t.item()), flush=True)


def local_decode():
    plan.decode(first_user=p2p.row(0), n_users=world, mean=True, out=out)


def direct_decode():
    plan.decode(n_users=world, mean=True, out=out, base_ptr=p2p.user0_record_ptr(), user_offsets=p2p.user_offsets())


timed("barrier", p2p.barrier)
timed("push (no barrier)", p2p.push)
timed("gather (no barrier)", p2p.gather)
timed("push + barrier", lambda: (p2p.push(), p2p.barrier()))
timed("barrier + gather", lambda: (p2p.barrier(), p2p.gather()))
timed("local decode", local_decode)
timed("pull-and-decode (no barrier)", direct_decode)
timed("push + barrier + local decode", lambda: (p2p.push(), p2p.barrier(), local_decode()))
timed("barrier + gather + local decode", lambda: (p2p.barrier(), p2p.gather(), local_decode()))
timed("barrier + pull-and-decode", lambda: (p2p.barrier(), direct_decode()))
dist.barrier()
p2p.close()
dist.destroy_process_group()
