import numpy as np
import torch

from .. import _lib
from ._common import (announce_dim, chunk_dim, device_of, load_codebook, single_segment,
                      uniforms_arg)
from .probabilistic_scalar_compressor import ProbabilisticScalarCompressor


class NearestNeighborCompressor(object):
    """HSQ: hyper-sphere vector quantization
    (reference compressors/nearest_neighbor_compressor.py:9-90).

    The gradient is viewed as [N/d, d]; every chunk is replaced by the index of
    the unit-norm codeword with the largest |inner product| and the signed
    projection u, which is quantized to n_bit by ProbabilisticScalarCompressor
    unless n_bit == 32.

    compress(vec)  -> [(lb, ub, l int32[N/d]), codes]   (or [u fp32[N/d], codes])
    decompress(sig) -> fp32 tensor of `shape`
    codes is uint8 when k_bit <= 8 else int32, as in the reference (:57).
    Codes, u, lb/ub and -- for the same uniforms -- l are bit-identical to the
    reference's CPU result.
    """

    def __init__(self, size, shape, args):
        c_dim, k_bit, n_bit = args.c_dim, args.k_bit, args.n_bit
        assert c_dim > 0
        assert k_bit >= 0
        assert n_bit > 0
        self.device = device_of(args)
        self.cuda = True
        self.size, self.shape = size, shape
        self.dim = chunk_dim(size, c_dim)
        announce_dim(c_dim, self.dim, size, shape)
        assert size % self.dim == 0, \
            "not divisible size {}  c_dim {} self.dim {}".format(size, c_dim, self.dim)
        self.K = self.dim if k_bit <= 0 else 2 ** k_bit
        if self.K == self.dim:
            # random orthogonal basis (reference :45-46)
            from scipy import stats
            codewords = stats.ortho_group.rvs(self.dim).astype(np.float32)
        else:
            codewords = load_codebook(self.dim, self.K)
        self.codewords = torch.from_numpy(np.ascontiguousarray(codewords)).to(self.device)
        self.code_dtype = torch.uint8 if k_bit <= 8 else torch.int32
        self.n_bit = n_bit
        self.compressed_norm = n_bit != 32
        self.random = args.random
        self.rng = getattr(args, "rng", "philox")
        self.algo = getattr(args, "hsq_algo", _lib.ALGO_AUTO)
        if self.compressed_norm:
            self.norm_compressor = ProbabilisticScalarCompressor(n_bit, args)
        self.n_chunks = size // self.dim
        self._seg = single_segment(self.n_chunks, self.device)
        self._ws_bytes = _lib.value("gq_hsq_encode_workspace_bytes", self.n_chunks, self.dim, self.K, 1)

    def compress(self, vec, uniforms=None):
        v = _lib.f32c(vec, "vec").reshape(-1)
        assert v.numel() == self.size
        dev = v.device
        n = self.n_chunks
        code_bytes = 1 if self.code_dtype == torch.uint8 else 4
        codes = torch.empty(n, dtype=self.code_dtype, device=dev)
        u = torch.empty(n, dtype=torch.float32, device=dev)
        ws = torch.empty(max(self._ws_bytes, 256), dtype=torch.uint8, device=dev)
        if not self.compressed_norm:
            _lib.call("gq_hsq_encode", _lib.ptr(v), n, self.dim, _lib.ptr(self.codewords), self.K,
                      _lib.ptr(self._seg), 1, 32, 0, None, 0, 0, _lib.ptr(codes), code_bytes, None, 4,
                      None, _lib.ptr(u), _lib.ptr(ws), ws.numel(), self.algo, _lib.stream())
            return [u, codes]
        random = 1 if self.random else 0
        r = uniforms_arg(uniforms, n, dev) if random else None
        if random and r is None and self.rng == "torch":
            # reference-faithful RNG: search first, then let the scalar compressor draw
            # from the CPU generator exactly when the reference would.
            _lib.call("gq_hsq_encode", _lib.ptr(v), n, self.dim, _lib.ptr(self.codewords), self.K,
                      _lib.ptr(self._seg), 1, 32, 0, None, 0, 0, _lib.ptr(codes), code_bytes, None, 4,
                      None, _lib.ptr(u), _lib.ptr(ws), ws.numel(), self.algo, _lib.stream())
            return [self.norm_compressor.compress(u), codes]
        seed, off = _lib.PHILOX.take(n) if (random and r is None) else (0, 0)
        l = torch.empty(n, dtype=torch.int32, device=dev)
        lbub = torch.empty(2, dtype=torch.float32, device=dev)
        _lib.call("gq_hsq_encode", _lib.ptr(v), n, self.dim, _lib.ptr(self.codewords), self.K,
                  _lib.ptr(self._seg), 1, self.n_bit, random, _lib.ptr(r), seed, off, _lib.ptr(codes),
                  code_bytes, _lib.ptr(l), 4, _lib.ptr(lbub), _lib.ptr(u), _lib.ptr(ws), ws.numel(),
                  self.algo, _lib.stream())
        return [(lbub[0], lbub[1], l), codes]

    def decompress(self, signature):
        norms, codes = signature
        codes = _lib.require_cuda(codes, "codes").contiguous().view(-1)
        dev = codes.device
        n = codes.numel()
        if codes.dtype == torch.uint8:
            code_bytes = 1
        else:
            code_bytes = 4
            codes = codes.to(torch.int32)
        out = torch.empty(n * self.dim, dtype=torch.float32, device=dev)
        if self.compressed_norm:
            lb, ub, l = norms
            l = l.contiguous().view(-1)
            l_bytes = 1 if l.dtype == torch.uint8 else 4
            if l_bytes == 4:
                l = l.to(torch.int32)
            lbub = torch.stack([lb.reshape(()), ub.reshape(())]).to(dev, torch.float32)
            _lib.call("gq_hsq_decode_reduce", _lib.ptr(codes), code_bytes, _lib.ptr(l), l_bytes,
                      _lib.ptr(lbub), None, 0, 1, n, self.dim, _lib.ptr(self.codewords), self.K,
                      _lib.ptr(self._seg), 1, self.n_bit, 0, 0, _lib.ptr(out), _lib.stream())
        else:
            nf = _lib.f32c(norms, "norms").reshape(-1)
            _lib.call("gq_hsq_decode_reduce", _lib.ptr(codes), code_bytes, None, 1, None, _lib.ptr(nf),
                      0, 1, n, self.dim, _lib.ptr(self.codewords), self.K, _lib.ptr(self._seg), 1, 32,
                      0, 0, _lib.ptr(out), _lib.stream())
        return out.view(self.shape)
