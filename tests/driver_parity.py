#!/usr/bin/env python
"""Loss parity of the training driver against the UNMODIFIED reference.

    python tests/driver_parity.py [--iters 300] [--quantizer hsq] [--users 4] [--mode ps] [--ef]

Runs the same FCN / synthetic-MNIST-shape job twice, in lock step:
  (a) the reference's quantizers + compressors (baseline/_ref, staged by baseline/make_ref.py) on
      the CPU, driven by the reference's one_iter flow (main.py:216-233);
  (b) main_b200.py's one_iter on cuda:0 with `--rng torch`, i.e. the B200 compressors consuming the
      reference's own CPU uniform stream (torch.manual_seed(seed); torch.rand in reference call order).
Same seed => same initial weights, same batches, same stochastic-rounding draws; the only difference
left is fp32 summation order in the model's forward / backward (CPU MKL vs cuBLAS).  Prints the two
loss curves side by side and their largest difference; exits non-zero beyond --tol.
"""
import argparse
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.path.join(ROOT, "baseline", "_ref")

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=300)
ap.add_argument("--quantizer", default="hsq")
ap.add_argument("--users", type=int, default=4)
ap.add_argument("--mode", default="ps")
ap.add_argument("--ef", action="store_true")
ap.add_argument("--two-phase", action="store_true")
ap.add_argument("--tol", type=float, default=0.05)
a = ap.parse_args()

import main_b200 as M  # noqa: E402

argv = ["--network", "fcn", "--dataset", "synthetic", "--quantizer", a.quantizer, "--mode", a.mode, "--c-dim", "16",
        "--k-bit", "8", "--n-bit", "6", "--cr", "100", "--num-users", str(a.users), "--logdir", "/tmp/gq_parity",
        "--max-iters", str(a.iters), "--train-size", "20000", "--test-size", "2000", "--rng", "torch"]
if a.ef:
    argv.append("--ef")
if a.two_phase:
    argv.append("--two-phase")
if a.quantizer == "qsgd":
    argv[argv.index("--c-dim") + 1] = "128"
    argv[argv.index("--n-bit") + 1] = "2"


def run_reference():
    if not os.path.isdir(os.path.join(REF, "compressors")):
        raise SystemExit("baseline/_ref is not staged (python baseline/make_ref.py)")
    args = M.parse_args(argv)
    args.no_cuda = True
    cwd = os.getcwd()
    os.chdir(REF)
    sys.path.insert(0, REF)
    try:
        import compressors as RC
        from quantizers import Quantizer as RQ
        comp = {'sgd': RC.IdenticalCompressor, 'qsgd': RC.QSGDCompressor, 'hsq': RC.NearestNeighborCompressor,
                'sign': RC.SignSGDCompressor, 'topk': RC.TopKSparsificationCompressor}[args.quantizer]
        torch.manual_seed(args.seed)
        train_loader, _ = M.make_loaders(args)
        model = M.FCN(num_classes=args.num_classes)
        q = RQ(comp, model.parameters(), args)
        lr, mom, wd = (1e-3, 0.0, 0.1) if args.quantizer == "sign" else (0.1, args.momentum, args.weight_decay)
        opt = torch.optim.SGD(model.parameters(), lr=lr, momentum=mom, weight_decay=wd)
        losses, it, epoch = [], 0, 1
        while it < a.iters:
            for data, target in train_loader:
                ub = len(data) // args.num_users
                td = [(data[u * ub:(u + 1) * ub], target[u * ub:(u + 1) * ub]) for u in range(args.num_users - 1)]
                td.append((data[(args.num_users - 1) * ub:], target[(args.num_users - 1) * ub:]))
                model.train()
                ls = []
                for u in range(args.num_users):            # main.py:221-230
                    opt.zero_grad()
                    loss = M.LOSS_FUNC(model(td[u][0]), td[u][1])
                    ls.append(loss)
                    loss.backward()
                    q.record(u, epoch=epoch)
                q.apply()
                opt.step()
                losses.append(torch.stack(ls).mean().item())
                it += 1
                if it >= a.iters:
                    break
            epoch += 1
        return losses
    finally:
        os.chdir(cwd)
        sys.path.remove(REF)
        for k in [k for k in sys.modules if k.split(".")[0] in ("compressors", "quantizers", "utils")]:
            del sys.modules[k]


torch.set_num_threads(os.cpu_count() or 1)
ref = run_reference()
state = M.main(argv)
ours = state["losses"]
n = min(len(ref), len(ours))
diff = np.abs(np.array(ref[:n]) - np.array(ours[:n]))
for i in list(range(0, n, max(n // 15, 1))) + [n - 1]:
    print("iter %4d  reference(CPU) %.6f   b200 %.6f   |diff| %.2e" % (i, ref[i], ours[i], diff[i]))
print("PARITY %s %s users=%d ef=%d two_phase=%d: %d iterations, first loss %.4f -> last %.4f (ref) / %.4f (b200), "
      "max |diff| %.3e, mean |diff| %.3e" % (a.quantizer, a.mode, a.users, int(a.ef), int(a.two_phase), n, ref[0], ref[n - 1],
                                            ours[n - 1], diff.max(), diff.mean()))
sys.exit(0 if diff.max() <= a.tol else 1)
