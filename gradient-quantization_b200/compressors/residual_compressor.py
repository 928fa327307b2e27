import torch

from .. import _lib
from .nearest_neighbor_compressor import NearestNeighborCompressor
from .probabilistic_vector_compressor import ProbabilisticVectorCompressor


class ResidualCompressor(object):
    """Two-stage residual quantizer (reference compressors/residual_compressor.py:7-32):
    stage 1 HSQ, stage 2 the probabilistic vector compressor on what stage 1 left.
    compress -> [signature_stage1, signature_stage2]; decompress -> sum of stages."""

    def __init__(self, size, shape, args):
        self.compressors = [
            NearestNeighborCompressor(size, shape, args),
            ProbabilisticVectorCompressor(size, shape, args),
        ]

    def compress(self, vec, uniforms=None):
        """uniforms: optional list of per-stage keyword dicts for deterministic checks."""
        residuals = _lib.f32c(vec, "vec").clone()
        flat = residuals.view(-1)
        signatures = []
        for i, compressor in enumerate(self.compressors):
            kw = uniforms[i] if uniforms is not None else {}
            signature = compressor.compress(residuals, **kw)
            decompressed = _lib.f32c(compressor.decompress(signature)).view(-1)
            _lib.call("gq_sub", _lib.ptr(flat), _lib.ptr(decompressed), flat.numel(), _lib.ptr(flat),
                      _lib.stream())
            signatures.append(signature)
        return signatures

    def decompress(self, signatures):
        parts = [c.decompress(s) for s, c in zip(signatures, self.compressors)]
        out = torch.empty_like(parts[0])
        _lib.call("gq_axpy", _lib.ptr(parts[0].contiguous()), _lib.ptr(parts[1].contiguous()), 1.0,
                  out.numel(), _lib.ptr(out), _lib.stream())
        return out
