import torch

from .. import _lib
from ._common import device_of


class TopKSparsificationCompressor(object):
    """Top-k sparsification (reference compressors/topk_sparsification_compressor.py:9-26):
    keep the k = size // args.cr entries of largest magnitude, zero the rest, return
    the dense tensor [1, size]; decompress views it back to `shape`.  Ties at the
    cut are resolved towards the lowest index (torch.topk leaves this open)."""

    def __init__(self, size, shape, args):
        self.device = device_of(args)
        self.cuda = True
        self.size, self.shape = size, shape
        self.users = 1
        self.k = size // args.cr
        self._seg = torch.tensor([0, size], dtype=torch.int64, device=self.device)
        self._k = torch.tensor([self.k], dtype=torch.int64, device=self.device)
        self._kp = torch.zeros(1, dtype=torch.int64, device=self.device)
        self._ws_bytes = _lib.value("gq_topk_workspace_bytes", size, 1)

    def compress(self, vec):
        v = _lib.f32c(vec, "vec").reshape(-1)
        assert v.numel() == self.size
        out = torch.empty_like(v)
        ws = torch.empty(self._ws_bytes, dtype=torch.uint8, device=v.device)
        _lib.call("gq_topk_select", _lib.ptr(v), self.size, _lib.ptr(self._seg), _lib.ptr(self._k),
                  _lib.ptr(self._kp), 1, _lib.ptr(out), None, None, _lib.ptr(ws), ws.numel(),
                  _lib.stream())
        return out.view(self.users, -1)

    def decompress(self, signature):
        return signature.view(self.shape)
