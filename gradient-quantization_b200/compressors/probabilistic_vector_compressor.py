import numpy as np
import torch

from .. import _lib
from ._common import device_of, load_codebook, single_segment, uniforms_arg
from .probabilistic_scalar_compressor import ProbabilisticScalarCompressor


class ProbabilisticVectorCompressor(object):
    """Unbiased probabilistic vector quantizer
    (reference compressors/probabilistic_vector_compressor.py:8-77).

    The shipped reference class cannot run (it opens './codebook/...', a directory
    that does not exist, and calls argmin on a bool tensor).  This
    class implements the algorithm those lines describe: p = pinv(C^T) v, pick
    codeword k with probability |p_k| / ||p||_1 by inverse-CDF sampling (first k
    whose cumulative probability reaches r - 1e-5), keep u = sign(p_k) ||p||_1, so
    that E[decompress] = v.  The codebook comes from codebooks/learned_codebook
    like HSQ's.  Parity with the reference is unpinned for this class.
    """

    def __init__(self, size, shape, args):
        c_dim, k_bit, n_bit = args.c_dim, args.k_bit, args.n_bit
        assert c_dim > 0
        assert k_bit > 0
        assert n_bit > 0
        self.device = device_of(args)
        self.cuda = True
        self.size, self.shape = size, shape
        self.dim = c_dim if c_dim < size else size
        assert size % self.dim == 0, "not divisible size {} dim {}".format(size, self.dim)
        self.K = 2 ** k_bit
        if self.K == self.dim:
            from scipy import stats
            codewords = stats.ortho_group.rvs(self.dim).astype(np.float32)
        else:
            codewords = load_codebook(self.dim, self.K)
        c_dagger = np.linalg.pinv(codewords.T).astype(np.float32)
        self.codewords = torch.from_numpy(np.ascontiguousarray(codewords)).to(self.device)
        self.c_dagger = torch.from_numpy(np.ascontiguousarray(c_dagger)).to(self.device)
        self.code_dtype = torch.uint8 if k_bit <= 8 else torch.int32
        self.n_bit = n_bit
        self.compressed_norm = n_bit != 32
        self.rng = getattr(args, "rng", "philox")
        if self.compressed_norm:
            self.norm_compressor = ProbabilisticScalarCompressor(n_bit, args)
        self.n_chunks = size // self.dim
        self._seg = single_segment(self.n_chunks, self.device)

    def compress(self, vec, uniforms=None, norm_uniforms=None):
        v = _lib.f32c(vec, "vec").reshape(-1)
        assert v.numel() == self.size
        dev = v.device
        n = self.n_chunks
        code_bytes = 1 if self.code_dtype == torch.uint8 else 4
        codes = torch.empty(n, dtype=self.code_dtype, device=dev)
        u = torch.empty(n, dtype=torch.float32, device=dev)
        r = uniforms_arg(uniforms, n, dev)
        if r is None and self.rng == "torch":
            r = torch.rand(n).to(dev)  # reference :52-54
        seed, off = _lib.PHILOX.take(n) if r is None else (0, 0)
        _lib.call("gq_pvc_search", _lib.ptr(v), n, self.dim, _lib.ptr(self.c_dagger), self.K, _lib.ptr(r),
                  seed, off, _lib.ptr(codes), code_bytes, _lib.ptr(u), _lib.stream())
        if self.compressed_norm:
            u = self.norm_compressor.compress(u, uniforms=norm_uniforms)
        return [u, codes]

    def decompress(self, signature):
        norms, codes = signature
        if self.compressed_norm:
            norms = self.norm_compressor.decompress(norms)
        codes = _lib.require_cuda(codes, "codes").contiguous().view(-1)
        n = codes.numel()
        code_bytes = 1 if codes.dtype == torch.uint8 else 4
        if code_bytes == 4:
            codes = codes.to(torch.int32)
        nf = _lib.f32c(norms, "norms").reshape(-1)
        out = torch.empty(n * self.dim, dtype=torch.float32, device=codes.device)
        _lib.call("gq_hsq_decode_reduce", _lib.ptr(codes), code_bytes, None, 1, None, _lib.ptr(nf), 0, 1,
                  n, self.dim, _lib.ptr(self.codewords), self.K, _lib.ptr(self._seg), 1, 32, 0, 0,
                  _lib.ptr(out), _lib.stream())
        return out.view(self.shape)
