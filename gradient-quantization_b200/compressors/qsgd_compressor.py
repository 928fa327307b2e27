import torch

from .. import _lib
from ._common import announce_dim, chunk_dim, device_of, uniforms_arg


class QSGDCompressor(object):
    """QSGD / TernGrad (reference compressors/qsgd_compressor.py:4-71): per chunk of
    `dim` elements an L-infinity norm, 2^n_bit stochastic levels and a sign bit.
    c_dim == 0 makes the whole tensor one chunk (TernGrad with n_bit == 1).

    compress(vec)  -> [norm fp32[M,1], signs bool[shape], l int32[shape]]
    decompress(sig) -> fp32 tensor of `shape`
    """

    def __init__(self, size, shape, args):
        self.random = args.random
        self.bit = args.n_bit
        c_dim = args.c_dim
        assert self.bit > 0
        self.device = device_of(args)
        self.cuda = True
        self.s = 2 ** self.bit
        self.size, self.shape = size, shape
        self.dim = chunk_dim(size, c_dim)
        announce_dim(c_dim, self.dim, size, shape)
        assert self.dim != 0, \
            "0 sub dimension size {}  c_dim {} self.dim {}".format(size, c_dim, self.dim)
        assert size % self.dim == 0, \
            "not divisible size {}  c_dim {} self.dim {}".format(size, c_dim, self.dim)
        self.M = size // self.dim
        self.code_dtype = torch.int32
        self.rng = getattr(args, "rng", "philox")

    def compress(self, vec, uniforms=None):
        v = _lib.f32c(vec, "vec").reshape(-1)
        assert v.numel() == self.size
        dev = v.device
        n = self.size
        norm = torch.empty(self.M, dtype=torch.float32, device=dev)
        signs = torch.empty(n, dtype=torch.uint8, device=dev)
        l = torch.empty(n, dtype=torch.int32, device=dev)
        random = 1 if self.random else 0
        r = uniforms_arg(uniforms, n, dev) if random else None
        if random and r is None and self.rng == "torch":
            r = torch.rand(n).to(dev)  # the reference's CPU draw (:58-60)
        seed, off = _lib.PHILOX.take(n) if (random and r is None) else (0, 0)
        _lib.call("gq_qsgd_encode", _lib.ptr(v), n, None, self.M, self.dim, self.bit, random, _lib.ptr(r),
                  seed, off, _lib.ptr(norm), _lib.ptr(signs), _lib.ptr(l), None, _lib.stream())
        return [norm.view(self.M, 1), signs.view(torch.bool).view(self.shape), l.view(self.shape)]

    def decompress(self, signature):
        norm, signs, l = signature
        assert l.shape == signs.shape
        dev = l.device
        norm = _lib.f32c(norm, "norm").reshape(-1)
        sg = signs.contiguous().view(-1)
        sg = sg.view(torch.uint8) if sg.dtype == torch.bool else sg.to(torch.uint8)
        lc = l.contiguous().view(-1).to(torch.int32)
        out = torch.empty(self.size, dtype=torch.float32, device=dev)
        _lib.call("gq_qsgd_decode_unpacked", _lib.ptr(norm), _lib.ptr(sg), _lib.ptr(lc), self.size, None,
                  self.M, self.dim, self.bit, _lib.ptr(out), _lib.stream())
        return out.view(self.shape)
