#!/usr/bin/env python
"""Times the UNMODIFIED reference (baseline/_ref, staged by make_ref.py) on the host CPU.

    python baseline/ref_runner.py --codec hsq --users 1 --steps 3 --warmup 1 [--workload resnet50|fcn|flat]

One step = quantizer.record(user) for every simulated user + quantizer.apply()
(quantizers/ps_quantizer.py:27-65 or ring_quantizer.py:25-49) over the workload's gradient tensors,
exactly as main.one_iter drives it (main.py:221-231) minus the model's forward/backward.
Runs in its own process (bench.py spawns it) so that the reference's module names (`compressors`,
`quantizers`, `utils`), its cwd-relative codebook path and the thread settings stay isolated.
Prints one JSON line.
"""
import argparse
import io
import json
import os
import sys
import time
from contextlib import redirect_stdout
from types import SimpleNamespace

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def shapes_of(workload):
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tests"))
    import util
    if workload == "resnet50":
        return util.resnet50_shapes()
    if workload == "fcn":
        return list(util.FCN_SHAPES)
    if workload == "flat":
        return [(25_600_000,)]
    raise ValueError(workload)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--codec", default="hsq", choices=["hsq", "qsgd", "terngrad", "sign", "topk"])
    ap.add_argument("--mode", default="ps", choices=["ps", "ring"])
    ap.add_argument("--users", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--workload", default="resnet50")
    ap.add_argument("--c-dim", type=int, default=None)
    ap.add_argument("--k-bit", type=int, default=8)
    ap.add_argument("--n-bit", type=int, default=None)
    ap.add_argument("--cr", type=int, default=100)
    ap.add_argument("--max-elems", type=int, default=0, help="keep only the first tensors up to this many elements")
    a = ap.parse_args()
    if not os.path.isdir(os.path.join(REF, "compressors")):
        print(json.dumps({"unavailable": "baseline/_ref is not staged (run baseline/make_ref.py where /root/reference exists)"}))
        return
    import numpy as np
    import torch
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    shapes = shapes_of(a.workload)
    if a.max_elems:
        kept, tot = [], 0
        for s in shapes:
            n = int(np.prod(s))
            if kept and tot + n > a.max_elems:
                continue
            kept.append(s)
            tot += n
        shapes = kept
    os.chdir(REF)                      # './codebooks/learned_codebook/...' (nearest_neighbor_compressor.py:50)
    sys.path.insert(0, REF)
    import compressors as C
    from quantizers import Quantizer
    cls = {"hsq": C.NearestNeighborCompressor, "qsgd": C.QSGDCompressor, "terngrad": C.QSGDCompressor,
           "sign": C.SignSGDCompressor, "topk": C.TopKSparsificationCompressor}[a.codec]
    c_dim = a.c_dim if a.c_dim is not None else {"hsq": 16, "qsgd": 128, "terngrad": 0}.get(a.codec, 16)
    n_bit = a.n_bit if a.n_bit is not None else {"hsq": 6, "qsgd": 2, "terngrad": 1}.get(a.codec, 6)
    args = SimpleNamespace(c_dim=c_dim, k_bit=a.k_bit, n_bit=n_bit, no_cuda=True, random=True, cr=a.cr, ef=False,
                           two_phase=False, mode=a.mode, scale="exp", num_users=a.users)
    params = [torch.nn.Parameter(torch.zeros(s)) for s in shapes]
    with redirect_stdout(io.StringIO()):       # the constructors print "alternate dimension ..." lines
        q = Quantizer(cls, params, args)
    gen = torch.Generator().manual_seed(1)
    grads = [[torch.randn(s, generator=gen) * 0.01 for s in shapes] for _ in range(a.users)]
    elems = sum(int(np.prod(s)) for s in shapes)

    def step():
        for u in range(a.users):
            for p, g in zip(params, grads[u]):
                p.grad = g.clone() if a.mode == "ring" else g   # ring adds into param.grad in place
            q.record(u, epoch=1)
        q.apply()

    for _ in range(a.warmup):
        step()
    times = []
    for _ in range(max(a.steps, 1)):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    t = float(np.mean(times))
    print(json.dumps({"value": elems * a.users / t, "unit": "elements/s", "seconds_per_step": t, "steps": len(times),
                      "elements_per_user": elems, "users": a.users, "cores": threads,
                      "torch_threads": torch.get_num_threads(), "torch": torch.__version__, "kind": "reference",
                      "codec": a.codec, "mode": a.mode, "workload": a.workload}))


if __name__ == "__main__":
    main()
