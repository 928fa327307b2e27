"""Helpers shared by the compressor classes."""
import os

import numpy as np
import torch

from .. import _lib
from ..utils.vecs_io import fvecs_read
from ..utils.vec_np import normalize

_PKG_CODEBOOKS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "codebooks")


def chunk_dim(size, c_dim):
    """The chunk-dimension rule shared by HSQ and QSGD
    (nearest_neighbor_compressor.py:23-29, qsgd_compressor.py:16-22): whole tensor
    when c_dim == 0 or the tensor is smaller than c_dim, else c_dim grown by
    x1.5 (dim // 2 * 3) up to ten times until it divides the size."""
    if c_dim == 0 or size < c_dim:
        return size
    dim = c_dim
    for _ in range(10):
        if size % dim != 0:
            dim = dim // 2 * 3
    return dim


def announce_dim(c_dim, dim, size, shape):
    if c_dim != dim:
        print("alternate dimension form {} to {}, size {} shape {}".format(c_dim, dim, size, shape))


def codebook_path(dim, K, subdir="learned_codebook"):
    """The reference opens './codebooks/learned_codebook/angular_dim_{d}_Ks_{K}.fvecs'
    relative to the cwd (nearest_neighbor_compressor.py:50-51).  Keep that lookup
    first (drop-in), then fall back to $GQ_CODEBOOK_DIR and the copies shipped
    inside this package."""
    name = "angular_dim_{}_Ks_{}.fvecs".format(dim, K)
    candidates = [os.path.join(".", "codebooks", subdir, name)]
    if os.environ.get("GQ_CODEBOOK_DIR"):
        candidates.append(os.path.join(os.environ["GQ_CODEBOOK_DIR"], subdir, name))
        candidates.append(os.path.join(os.environ["GQ_CODEBOOK_DIR"], name))
    candidates.append(os.path.join(_PKG_CODEBOOKS, subdir, name))
    for c in candidates:
        if os.path.exists(c):
            return c
    raise FileNotFoundError("no codebook file for dim={} K={} (looked in: {})".format(dim, K, ", ".join(candidates)))


_CODEBOOK_CACHE = {}


def load_codebook(dim, K):
    """Unit-norm fp32 codebook [K, dim] as a numpy array (cached per file)."""
    path = os.path.abspath(codebook_path(dim, K))
    if path not in _CODEBOOK_CACHE:
        _CODEBOOK_CACHE[path] = np.ascontiguousarray(normalize(fvecs_read(path))[1], dtype=np.float32)
    return _CODEBOOK_CACHE[path]


def device_of(args):
    """The reference runs on CPU when args.no_cuda is set; this implementation
    has no CPU path and says so instead of silently doing something else."""
    if getattr(args, "no_cuda", False):
        raise _lib.GQError("args.no_cuda is set, but gradient-quantization_b200 is CUDA-only "
                           "(no CPU fallback by design)")
    if not torch.cuda.is_available():
        raise _lib.GQError("no CUDA device available; gradient-quantization_b200 is CUDA-only")
    return torch.device("cuda", torch.cuda.current_device())


def single_segment(n, device):
    return torch.tensor([0, n], dtype=torch.int64, device=device)


def uniforms_arg(uniforms, n, device):
    """Optional externally supplied U[0,1) draws (deterministic parity checks)."""
    if uniforms is None:
        return None
    u = torch.as_tensor(uniforms, dtype=torch.float32)
    if u.numel() < n:
        raise ValueError("need {} uniforms, got {}".format(n, u.numel()))
    return u.reshape(-1)[:n].to(device).contiguous()
