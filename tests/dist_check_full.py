#!/usr/bin/env python
"""Multi-GPU parity at BASELINE's full size (run under torchrun, one user per GPU):
ps record/apply of the ResNet-50 gradient over the packed-record exchange -- by default the push
fused into the encode kernel + the delivery-flag wait inside the decode kernel -- for several
consecutive steps (both parities of the double-buffered receive blocks, epoch counting), against
the oracle's single-process result (rank 0) and bit-identical across ranks.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29512 tests/dist_check_full.py [--steps 4]
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gq_b200  # noqa: E402
from oracle import gq_oracle as O  # noqa: E402
from util import codebook, make_args, resnet50_shapes  # noqa: E402

steps = int(sys.argv[sys.argv.index("--steps") + 1]) if "--steps" in sys.argv else 4
rank = int(os.environ["RANK"])
world = int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)

shapes = resnet50_shapes()
sizes = [int(np.prod(s)) for s in shapes]
a = make_args(mode="ps", num_users=world)
params = [torch.nn.Parameter(torch.zeros(s, device=dev)) for s in shapes]
q = gq_b200.Quantizer(gq_b200.NearestNeighborCompressor, params, a)
plan = q.plan
per_user = sum(n // 16 for n in sizes if n > 1000)
cb = codebook(16, 256)
codecs = [O.HSQ(n, s, cb, 6, True) if n > 1000 else O.Identity() for n, s in zip(sizes, shapes)]
fails = 0
for step in range(steps):
    # every rank can regenerate every user's gradient and uniforms (seeded by step and user)
    def user_data(u):
        g = torch.Generator(device=dev)
        g.manual_seed(1000 * step + u)
        flat = torch.randn(plan.arena_elems, device=dev, generator=g) * (0.01 * (1 + u))
        r = np.random.RandomState(7000 + 100 * step + u).random_sample(per_user).astype(np.float32)
        return flat, r
    flat, r = user_data(rank)
    for p, v in zip(params, plan.views(flat)):
        p.grad = v.clone()
    parts, used = plan.split_uniform_stream(r)
    q.record(rank, epoch=1, uniforms=parts)
    q.apply()
    torch.cuda.synchronize()
    mine = torch.cat([p.grad.data.reshape(-1) for p in params])
    ref0 = mine.clone()
    dist.broadcast(ref0, src=0)
    same = torch.equal(mine, ref0)
    ok = same
    if rank == 0:
        grads, draws = [], []
        for u in range(world):
            fu, ru = user_data(u)
            grads.append([v.cpu().numpy() for v in plan.views(fu)])
            draws.append(ru)
        ref = O.ps_step(codecs, grads, O.UniformStream(np.concatenate(draws)))
        for i, (p, r_) in enumerate(zip(params, ref)):
            got = p.grad.data.cpu().numpy()
            r_ = r_.reshape(got.shape)
            if got.size >= 256:
                ok = ok and np.array_equal(got, r_)
            else:
                ok = ok and bool(np.abs(got - r_).max() <= 1e-6 * max(np.abs(r_).max(), 1e-30))
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    if rank == 0:
        print("full-size ps hsq world=%d step %d (%s): %s" % (world, step, q.exchange_name(),
              "OK (oracle-exact on rank 0, identical on every rank)" if flag.item() == 0 else "MISMATCH"), flush=True)
    fails += int(flag.item())
dist.barrier()
dist.destroy_process_group()
sys.exit(1 if fails else 0)
