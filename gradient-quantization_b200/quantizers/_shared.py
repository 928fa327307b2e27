"""Pieces shared by the ps and ring quantizers."""
import math

import torch
import torch.distributed as dist

from .. import _lib
from ..compressors import IdenticalCompressor
from .fused import FusedPlan


def feedback_scale(args, epoch):
    """scale of the error-feedback term (ps_quantizer.py:28-31, ring_quantizer.py:26-29)."""
    if args.scale == 'exp':
        return 2 / (math.exp(-epoch) + 1) - 1
    return float(args.scale)


def dist_world():
    """(rank, world) when torch.distributed is up with more than one rank, else (0, 1)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


class QuantizerBase(object):
    """Common construction: one compressor per parameter (size > 1000, else identity;
    ps_quantizer.py:15-20), error-feedback buffers attached to the parameters
    (:21-25), and -- when the compressor class allows it -- the fused whole-model
    plan that the hot path actually runs on."""

    def __init__(self, Compressor, parameters, args):
        self.parameters = list(parameters)
        self.num_layers = len(self.parameters)
        self.args = args
        self.error_feedback = args.ef
        self.compressed_gradients = [list() for _ in range(self.num_layers)]
        if self.num_layers and not self.parameters[0].is_cuda:
            raise _lib.GQError("parameters must live on a CUDA device (no CPU path)")
        self.device = self.parameters[0].device if self.num_layers else torch.device("cuda")
        self.rank, self.world = dist_world()
        self.distributed = self.world > 1
        if self.distributed and args.num_users != self.world:
            raise _lib.GQError("distributed mode maps one user per rank: num_users (%d) != world size (%d)"
                               % (args.num_users, self.world))
        use_fused = getattr(args, "fused", True) and FusedPlan.supports(Compressor)
        if use_fused and Compressor.__name__ == "NearestNeighborCompressor":
            # K == dim asks for a per-tensor random orthogonal basis (nearest_neighbor_compressor.py:45):
            # per-parameter path only.  That is k_bit <= 0, or 2 ** k_bit equal to the (possibly
            # escalated) chunk dimension of any compressed tensor, e.g. c_dim = 16 with k_bit = 4.
            from ..compressors._common import chunk_dim
            sizes = [int(p.numel()) for p in self.parameters if int(p.numel()) > 1000]
            use_fused = args.k_bit > 0 and all(chunk_dim(n, args.c_dim) != 2 ** args.k_bit for n in sizes)
        if self.distributed and not use_fused:
            raise _lib.GQError("distributed exchange needs a compressor with a packed wire format")
        self.plan = None
        self.compressors = list()
        if use_fused:
            self.plan = FusedPlan(Compressor, [p.shape for p in self.parameters], args, self.device,
                                  args.num_users)
        for param in self.parameters:
            param_size = param.flatten().shape[0]
            if self.plan is None:
                self.compressors.append(
                    Compressor(param_size, param.shape, args) if param_size > 1000
                    else IdenticalCompressor())
            if self.error_feedback:
                param.error = [torch.zeros_like(param) for _ in range(args.num_users)]

    # helpers for the fused path -------------------------------------------
    def _grads(self):
        return [p.grad.data for p in self.parameters]

    def _set_grads_from(self, buf):
        """apply(): param.grad.data is REPLACED (ps_quantizer.py:63) by views of buf."""
        for p, v in zip(self.parameters, self.plan.views(buf)):
            p.grad.data = v

    def _ef_buffers(self, user):
        """Error-feedback state of one user as an arena-layout buffer (lazy)."""
        if not hasattr(self, "_ef_flat"):
            self._ef_flat = {}
        if user not in self._ef_flat:
            flat = torch.zeros_like(self.plan.arena)
            for p, v in zip(self.parameters, self.plan.views(flat)):
                v.copy_(p.error[user])
                p.error[user].data = v     # param.error[user] stays a live view
            self._ef_flat[user] = flat
        return self._ef_flat[user]
