import torch

from .. import _lib
from . import exchange as xch
from ._shared import QuantizerBase, feedback_scale


class PSQuantizer(QuantizerBase):
    """Parameter-server exchange (reference quantizers/ps_quantizer.py:6-65).

    record(user, epoch): compress that user's gradients (after adding
    scale * error[user] under --ef, and storing the new error).
    apply(): average the users' decompressed gradients, optionally compress the
    average once more (--two-phase), and put the result in param.grad.data.

    Fused path (HSQ, QSGD, sign, top-k, identity): record() packs the whole model
    into records[user]; apply() is one decode-and-average pass over all users.
    With torch.distributed initialised (one user per rank), record(rank) packs the
    local gradient and apply() all-gathers the packed records over NCCL first --
    the packed codes, never fp32 gradients, cross NVLink.
    """

    def __init__(self, Compressor, parameters, args):
        super().__init__(Compressor, parameters, args)
        self.two_phase = self.args.two_phase
        if self.error_feedback and self.two_phase:
            for param in self.parameters:
                param.server_error = torch.zeros_like(param)
        if self.plan is not None and self.two_phase:
            self._phase2_plan = None  # built lazily (a 1-user plan for the averaged gradient)

    # ------------------------------------------------------------------ record
    def record(self, user, epoch, uniforms=None):
        scale = feedback_scale(self.args, epoch)
        if self.plan is None:
            return self._record_per_parameter(user, scale)
        plan = self.plan
        if self.distributed and user != self.rank:
            raise _lib.GQError("distributed mode: rank %d records user %d only" % (self.rank, self.rank))
        plan.gather(self._grads())
        if self.error_feedback:
            err = self._ef_buffers(user)
            n = plan.arena.numel()
            # grad += scale * error[user]   (ps_quantizer.py:35)
            _lib.call("gq_axpy", _lib.ptr(plan.arena), _lib.ptr(err), float(scale), n,
                      _lib.ptr(plan.arena), _lib.stream())
            for p, v in zip(self.parameters, plan.views()):
                p.grad.data = v               # the reference mutates param.grad in place
            plan.encode(user, uniforms=uniforms)
            # error[user] = grad - decompress(compress(grad))   (:36-39)
            dec = plan.decode(first_user=user, n_users=1, mean=False, out=self._scratch())
            _lib.call("gq_sub", _lib.ptr(plan.arena), _lib.ptr(dec), n, _lib.ptr(err), _lib.stream())
        else:
            plan.encode(user, uniforms=uniforms)

    def _scratch(self):
        if not hasattr(self, "_scratch_buf"):
            self._scratch_buf = torch.empty_like(self.plan.arena)
        return self._scratch_buf

    def _record_per_parameter(self, user, scale):
        for i, param in enumerate(self.parameters):
            if self.error_feedback:
                param.grad.data.add_(scale * param.error[user])
                decompressed_g = self.compressors[i].decompress(
                    self.compressors[i].compress(param.grad.data))
                param.error[user].data = param.grad.data - decompressed_g
            else:
                decompressed_g = self.compressors[i].decompress(
                    self.compressors[i].compress(param.grad.data))
            self.compressed_gradients[i].append(decompressed_g)

    # ------------------------------------------------------------------- apply
    def exchange(self):
        """All-gather the packed records: rank r's record is already in slot r."""
        if self.distributed:
            xch.ps_all_gather(self.plan.records, self.rank)

    def apply(self, uniforms=None):
        if self.plan is None:
            return self._apply_per_parameter()
        plan = self.plan
        self.exchange()
        g = plan.decode(mean=True, out=plan.arena)
        if self.two_phase:
            g = self._second_phase(g, uniforms)
        self._set_grads_from(g)

    def _second_phase(self, g, uniforms):
        """Compress the averaged gradient once more (ps_quantizer.py:52-61).  Every rank
        holds the same average; with identical uniforms (or the shared Philox state)
        every rank computes the same result."""
        from .fused import FusedPlan
        if self._phase2_plan is None:
            self._phase2_plan = FusedPlan(self.plan.Compressor, self.plan.shapes, self.args, self.device, 1)
        p2 = self._phase2_plan
        n = g.numel()
        if self.error_feedback:
            if not hasattr(self, "_server_err"):
                self._server_err = torch.zeros_like(g)
                for p, v in zip(self.parameters, self.plan.views(self._server_err)):
                    p.server_error = v
            _lib.call("gq_axpy", _lib.ptr(g), _lib.ptr(self._server_err), 1.0, n, _lib.ptr(g), _lib.stream())
        p2.encode(0, src=g, uniforms=uniforms)
        dec = p2.decode(mean=False, out=self._scratch())
        if self.error_feedback:
            _lib.call("gq_sub", _lib.ptr(g), _lib.ptr(dec), n, _lib.ptr(self._server_err), _lib.stream())
        g.copy_(dec)
        return g

    def _apply_per_parameter(self):
        for i, param in enumerate(self.parameters):
            # stack(...).mean(0) with the reference's CPU semantics: sum in user order,
            # then a true division (torch's CUDA mean multiplies by 1/U instead)
            stacked = torch.stack([_lib.f32c(t) for t in self.compressed_gradients[i]], dim=0)
            g = torch.empty_like(stacked[0])
            _lib.call("gq_f32_reduce_users", _lib.ptr(stacked), g.numel() * 4, stacked.shape[0],
                      g.numel(), 1, 0, _lib.ptr(g), _lib.stream())
            if self.two_phase:
                if self.error_feedback:
                    g.add_(param.server_error)
                    decompressed_g = self.compressors[i].decompress(self.compressors[i].compress(g))
                    param.server_error = g - decompressed_g
                    g = decompressed_g
                else:
                    g = self.compressors[i].decompress(self.compressors[i].compress(g))
            param.grad.data = g
        for compressed in self.compressed_gradients:
            compressed.clear()
