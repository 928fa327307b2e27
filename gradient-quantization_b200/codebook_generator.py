"""Codebook generator: k-means codebooks for HSQ, on the GPU, for any (d, K).

The reference ships `codebook_generator.py` (codebook_generator.py:14-31): Euclidean k-means
(scipy `kmeans2`, minit='points', 20 iterations) on 1 M unit-normalised Gaussian samples, seed 808,
for d = 1..65 and K in {32, 64, 256, 512, 1024}, written as `angular_dim_{d}_Ks_{K}.fvecs`; its
`codebooks/learned_codebook/` additionally holds K up to 4096 -- but nothing for K = 2^16
(BASELINE.json config 4) nor for the escalated chunk dims 36 / 192 that the dim rule
(nearest_neighbor_compressor.py:27-29) can produce.  This module fills that gap:

    python -m gq_b200.codebook_generator --dim 16 --K 65536 --out codebooks/learned_codebook
    train_codebook(dimension, Ks, train_size=1000000, iter=20)      # the reference's signature
    generate(dims=..., Ks=..., out_dir=...)

Two objectives:
  * "euclid" (default, what the reference's generator optimises): Lloyd iterations on the unit sphere,
    assignment argmin ||x - c||^2 in row blocks (torch ops on whatever device holds the samples);
  * "hsq": the assignment HSQ itself performs -- argmax_k |<c_k, x>| with a signed projection, computed
    by this package's own search kernel (gq_hsq_search: tcgen05 for d = 16 / K = 256, the exact CUDA-core
    kernel otherwise); samples are folded onto their codeword's side (x * sign(u)) before the mean, so a
    codeword and its negation are one cluster, which is how compress() uses them.
Empty clusters are re-seeded from the samples with the largest quantisation error.  Deterministic
for a given seed and device type.  Offline tooling: not on the gradient hot path.
"""
import argparse
import os

import numpy as np
import torch

from .utils.vec_np import normalize
from .utils.vecs_io import fvecs_writer


def _samples(dimension, train_size, seed, device):
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    x = torch.randn(train_size, dimension, generator=g, dtype=torch.float32)
    x = x / x.norm(dim=1, keepdim=True).clamp_min(1e-30)
    return x.to(device)


def _assign_euclid(x, c, block=8192):
    """argmin ||x - c||^2 = argmax (<x, c> - ||c||^2 / 2); returns (index, squared distance)."""
    half = 0.5 * (c * c).sum(1)
    idx = torch.empty(x.shape[0], dtype=torch.int64, device=x.device)
    d2 = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
    rows = max(1, min(block, (1 << 28) // max(c.shape[0], 1)))
    for i in range(0, x.shape[0], rows):
        xb = x[i:i + rows]
        s = xb @ c.t() - half
        m, j = s.max(dim=1)
        idx[i:i + rows] = j
        d2[i:i + rows] = (xb * xb).sum(1) - 2.0 * m
    return idx, d2


def _assign_hsq(x, c):
    """HSQ's own assignment through the package's search kernel: (index, signed projection u)."""
    from . import _lib
    n, d = x.shape
    K = c.shape[0]
    code_bytes = 1 if K <= 256 else 4
    codes = torch.empty(n, dtype=torch.uint8 if code_bytes == 1 else torch.int32, device=x.device)
    u = torch.empty(n, dtype=torch.float32, device=x.device)
    seg = torch.tensor([0, n], dtype=torch.int64, device=x.device)
    ws = torch.empty(1 << 16, dtype=torch.uint8, device=x.device)
    cc = c.contiguous()
    _lib.call("gq_hsq_search", x.data_ptr(), n, d, cc.data_ptr(), K, codes.data_ptr(), code_bytes, u.data_ptr(),
              seg.data_ptr(), 1, None, ws.data_ptr(), ws.numel(), _lib.ALGO_AUTO, _lib.stream())
    return codes.to(torch.int64), u


def kmeans(x, Ks, iters=20, objective="euclid", seed=808, normalize_rows=False):
    """Lloyd iterations on the rows of x (unit vectors); returns the [Ks, d] centroids (fp32 tensor)."""
    n, d = x.shape
    if Ks > n:
        raise ValueError("more centroids (%d) than samples (%d)" % (Ks, n))
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed) + 1)
    c = x[torch.randperm(n, generator=g)[:Ks].to(x.device)].clone()      # minit='points'
    for _ in range(iters):
        if objective == "hsq":
            idx, u = _assign_hsq(x, c)
            sgn = torch.where(u < 0, -torch.ones_like(u), torch.ones_like(u))
            xs = x * sgn[:, None]
            err = 1.0 - u * u                                            # ||x - u c||^2 for unit x, c
        else:
            idx, err = _assign_euclid(x, c)
            xs = x
        sums = torch.zeros(Ks, d, dtype=torch.float32, device=x.device).index_add_(0, idx, xs)
        cnt = torch.zeros(Ks, dtype=torch.float32, device=x.device).index_add_(0, idx, torch.ones(n, device=x.device))
        new = sums / cnt.clamp_min(1.0)[:, None]
        empty = torch.nonzero(cnt == 0).flatten()
        if empty.numel():                                                # re-seed empty clusters from the worst-fitted samples
            worst = torch.topk(err, int(empty.numel())).indices
            new[empty] = x[worst]
        if objective == "hsq":                                           # the search assumes unit-norm codewords
            new = new / new.norm(dim=1, keepdim=True).clamp_min(1e-30)
        c = new
    if normalize_rows:
        c = c / c.norm(dim=1, keepdim=True).clamp_min(1e-30)
    return c


def train_codebook(dimension, Ks, train_size=1000000, iter=20, seed=808, objective="euclid", device=None):
    """Same call as the reference's train_codebook (codebook_generator.py:14-21): [Ks, dimension] fp32
    numpy array of raw centroids (the compressors normalise rows on load, like the reference)."""
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    if objective == "hsq" and torch.device(device).type != "cuda":
        raise RuntimeError("objective='hsq' runs the CUDA search kernel: it needs a CUDA device")
    x = _samples(dimension, max(train_size, Ks), seed, device)
    c = kmeans(x, Ks, iters=iter, objective=objective, seed=seed)
    return np.ascontiguousarray(c.cpu().numpy(), dtype=np.float32)


def generate(dims=range(1, 66), Ks=(32, 64, 256, 512, 1024), out_dir="codebook", train_size=1000000, iter=20,
             seed=808, objective="euclid", overwrite=False):
    """The reference's generate() (codebook_generator.py:23-31) for arbitrary dims / sizes."""
    os.makedirs(out_dir, exist_ok=True)
    written = []
    for dim in dims:
        for K in Ks:
            path = os.path.join(out_dir, "angular_dim_{}_Ks_{}.fvecs".format(dim, K))
            if os.path.exists(path):
                if not overwrite:
                    continue
                os.remove(path)                                          # fvecs_writer appends
            cb = train_codebook(dim, K, train_size=train_size, iter=iter, seed=seed + 131 * dim + K, objective=objective)
            assert cb.shape == (K, dim)
            fvecs_writer(path, cb)
            written.append(path)
            print("writing codebook into file {}".format(path))
    return written


def quantisation_error(codebook, n=200000, seed=1):
    """Mean squared HSQ error E[1 - max_k <c_k, x>^2] on fresh unit Gaussians (a quality figure)."""
    cb = torch.from_numpy(normalize(np.asarray(codebook, dtype=np.float32))[1])
    x = _samples(cb.shape[1], n, seed, cb.device)
    best = torch.zeros(n)
    for i in range(0, n, 8192):
        best[i:i + 8192] = (x[i:i + 8192] @ cb.t()).abs().max(dim=1).values
    return float((1.0 - best * best).mean())


def main():
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--dim", type=int, nargs="+", required=True)
    ap.add_argument("--K", type=int, nargs="+", required=True)
    ap.add_argument("--out", type=str, default=os.path.join("codebooks", "learned_codebook"))
    ap.add_argument("--train-size", type=int, default=1000000)
    ap.add_argument("--iter", type=int, default=20)
    ap.add_argument("--seed", type=int, default=808)
    ap.add_argument("--objective", choices=["euclid", "hsq"], default="euclid")
    ap.add_argument("--overwrite", action="store_true")
    a = ap.parse_args()
    generate(a.dim, a.K, a.out, a.train_size, a.iter, a.seed, a.objective, a.overwrite)


if __name__ == "__main__":
    main()
