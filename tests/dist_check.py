#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun, one user per GPU, NCCL):
ps and ring record/apply over the packed-record exchange must equal the oracle's
single-process result on all users' gradients, bit for bit.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/dist_check.py
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
from util import FCN_SHAPES, codebook, gen_input, make_args, torch_uniform_stream  # noqa: E402

rank = int(os.environ["RANK"])
world = int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)

shapes = FCN_SHAPES + [(64, 3, 3, 3)]
cases = (("ps", "hsq"), ("ring", "hsq"), ("ps", "qsgd"), ("ps", "sign"), ("ps", "topk"), ("ring", "qsgd"))
if os.environ.get("GQ_DIST_SHAPES") == "resnet50":      # the headline workload (the staged ring cuts it into parts)
    from util import resnet50_shapes
    shapes = resnet50_shapes()
if os.environ.get("GQ_DIST_CASES"):                     # e.g. "ring:hsq,ps:hsq"
    cases = tuple(tuple(c.split(":")) for c in os.environ["GQ_DIST_CASES"].split(","))
sizes = [int(np.prod(s)) for s in shapes]
fails = 0
for mode, quant in cases:
    a = make_args(mode=mode, num_users=world, c_dim=16 if quant == "hsq" else 128,
                  n_bit=6 if quant == "hsq" else 2, cr=100)
    Comp = {"hsq": gq_b200.NearestNeighborCompressor, "qsgd": gq_b200.QSGDCompressor,
            "sign": gq_b200.SignSGDCompressor, "topk": gq_b200.TopKSparsificationCompressor}[quant]
    params = [torch.nn.Parameter(torch.zeros(s, device=dev)) for s in shapes]
    q = gq_b200.Quantizer(Comp, params, a)
    assert q.distributed and q.rank == rank
    grads = [[gen_input(7000 + 10 * u + i, n).reshape(s) for i, (n, s) in enumerate(zip(sizes, shapes))]
             for u in range(world)]
    per_user = sum((n // 16 if quant == "hsq" else n) for n in sizes if n > 1000) if quant in ("hsq", "qsgd") else 0
    stream = torch_uniform_stream(99, per_user * world)
    for p, g in zip(params, grads[rank]):
        p.grad = torch.from_numpy(g).to(dev)
    parts, used = q.plan.split_uniform_stream(stream[rank * per_user:(rank + 1) * per_user])
    q.record(rank, epoch=1, uniforms=parts)
    q.apply()
    torch.cuda.synchronize()
    # oracle
    codecs = []
    for n, s in zip(sizes, shapes):
        if n <= 1000:
            codecs.append(O.Identity())
        elif quant == "hsq":
            codecs.append(O.HSQ(n, s, codebook(O.chunk_dim(n, 16), 256), 6, True))
        elif quant == "qsgd":
            codecs.append(O.QSGD(n, s, 128, 2, True))
        elif quant == "sign":
            codecs.append(O.Sign(n, s))
        else:
            codecs.append(O.TopK(n, s, 100))
    ostream = O.UniformStream(stream)
    ref = O.ps_step(codecs, grads, ostream) if mode == "ps" else O.ring_step(codecs, grads, ostream)
    ok = True
    for p, r in zip(params, ref):
        got = p.grad.data.cpu().numpy()
        if not np.array_equal(got, r.reshape(got.shape)):
            ok = False
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    if rank == 0:
        print("%-4s %-5s world=%d [%s]: %s" % (mode, quant, world, q.exchange_name()[:60], "OK (bit-exact on every rank)" if flag.item() == 0 else "MISMATCH"), flush=True)
    fails += int(flag.item())
dist.barrier()
dist.destroy_process_group()
sys.exit(1 if fails else 0)
