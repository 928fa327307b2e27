import torch

from .. import _lib
from . import exchange as xch
from ._shared import QuantizerBase, feedback_scale


class RingQuantizer(QuantizerBase):
    """Ring exchange (reference quantizers/ring_quantizer.py:7-49): a lossy running
    SUM passed from user to user, acc_r = D(C(g_r + acc_{r-1})); apply() hands the
    last user's decompressed sum (not divided by the number of users) to every
    parameter.

    Fused path: record(user) decodes records[user-1] straight into the gradient
    arena (decode-accumulate), then packs the sum into records[user].  With one
    user per rank the packed record travels rank r -> r+1 with NCCL send/recv and
    the last rank broadcasts its record; fp32 never crosses NVLink.
    """

    def __init__(self, Compressor, parameters, args):
        super().__init__(Compressor, parameters, args)
        # Pipelined chain: the chain is per tensor, so the plan is cut into stages and rank r
        # works on stage s while rank r + 1 works on stage s - 1: all ranks are busy after U - 1 stages.
        # args.ring_parts / GQ_RING_PARTS; default 4 stages from three ranks on, 2 for two ranks.
        import os
        self.parts = 1
        if self.plan is not None and not self.error_feedback:
            want = getattr(args, "ring_parts", None)
            if want is None:
                want = os.environ.get("GQ_RING_PARTS")
            if want is None:
                want = (4 if self.world >= 3 else 2) if self.distributed else 1
            self.parts = self.plan.make_parts(int(want))

    def record(self, user, epoch, uniforms=None):
        scale = feedback_scale(self.args, epoch)
        if self.plan is None:
            return self._record_per_parameter(user, scale)
        plan = self.plan
        if self.distributed and user != self.rank:
            raise _lib.GQError("distributed mode: rank %d records user %d only" % (self.rank, self.rank))
        plan.gather(self._grads())
        if self.parts > 1:
            return self._chain_parts(user, plan.arena, uniforms)
        if user != 0:
            if self.distributed:
                xch.ring_receive_previous(plan.records, user)
            # grad += previous hop's decompressed running sum   (ring_quantizer.py:31-32)
            plan.decode(first_user=user - 1, n_users=1, mean=False, accumulate=True, out=plan.arena)
        n = plan.arena.numel()
        if self.error_feedback and plan.supports_inplace_feedback():
            # E_u = (grad + previous) + scale * E_u in one multi-tensor pass over the arena's views
            # (same order of additions as ring_quantizer.py:32-34), encoded in place, and the decode of
            # the fresh record subtracts itself: E_u = grad' - decompress(compress(grad'))  (:36-38)
            err = self._ef_buffers(user)
            plan.gather(plan.views(), buf=err, feedback=1, scale=scale)
            plan.encode(user, src=err, uniforms=uniforms, rng_user=user)
            plan.decode(first_user=user, n_users=1, mean=False, accumulate=2, out=err)
        elif self.error_feedback:
            err = self._ef_buffers(user)
            _lib.call("gq_axpy", _lib.ptr(plan.arena), _lib.ptr(err), float(scale), n,
                      _lib.ptr(plan.arena), _lib.stream())
            plan.encode(user, uniforms=uniforms, rng_user=user)
            if not hasattr(self, "_scratch_buf"):
                self._scratch_buf = torch.empty_like(plan.arena)
            dec = plan.decode(first_user=user, n_users=1, mean=False, out=self._scratch_buf)
            _lib.call("gq_sub", _lib.ptr(plan.arena), _lib.ptr(dec), n, _lib.ptr(err), _lib.stream())
        else:
            plan.encode(user, uniforms=uniforms, rng_user=user)
        if self.distributed:
            xch.ring_send_next(plan.records, user, self.world)

    def _chain_parts(self, user, buf, uniforms=None):
        """This user's hop of the chain, stage by stage: receive the stage of the previous hop's record,
        add its decode to the local gradient, encode the sum, send the stage on."""
        plan = self.plan
        for p in range(self.parts):
            ranges = plan.part_byte_ranges(p)
            if user != 0:
                if self.distributed:
                    xch.ring_receive_previous_part(plan.records, user, ranges)
                plan.decode(first_user=user - 1, n_users=1, mean=False, accumulate=True, out=buf, part=p)
            plan.encode(user, src=buf, uniforms=uniforms, rng_user=user, part=p)
            if self.distributed:
                xch.ring_send_next_part(plan.records, user, self.world, ranges)

    def step_buffers(self, src, out):
        """One ring step on raw arena-shaped device buffers (bench.py): this rank's hop -- receive the
        previous hop's record, add its decode to the local gradient IN PLACE (ring_quantizer.py:31-32
        mutates param.grad), encode the sum, send it on -- then the broadcast of the last hop and the
        final decode into `out`.  Single process: all users' hops one after the other on `src`."""
        plan = self.plan
        users = [self.rank] if self.distributed else range(self.args.num_users)
        for user in users:
            if self.parts > 1:
                self._chain_parts(user, src)
                continue
            if user != 0:
                if self.distributed:
                    xch.ring_receive_previous(plan.records, user)
                plan.decode(first_user=user - 1, n_users=1, mean=False, accumulate=True, out=src)
            plan.encode(user, src=src)
            if self.distributed:
                xch.ring_send_next(plan.records, user, self.world)
        last = self.args.num_users - 1
        if self.distributed:
            xch.ring_broadcast_last(plan.records, self.world)
        return plan.decode(first_user=last, n_users=1, mean=False, out=out)

    def launches_per_step(self):
        """My kernel launches per rank and step: (decode-accumulate +) encode + final decode."""
        # every extra stage of the pipelined chain adds one encode (and one decode-accumulate) launch per
        # partitioned group
        extra = (self.parts - 1) * sum(1 for g in self.plan.groups if getattr(g, "part_cuts", None))
        enc = self.plan.launches_per_encode() + extra
        dec = self.plan.launches_per_decode(1)
        if self.distributed:
            return enc + dec + ((dec + extra) if self.rank > 0 else 0)
        u = self.args.num_users
        return u * enc + (u - 1) * (dec + extra) + dec

    def exchange_name(self):
        if not self.distributed:
            return "none"
        if self.parts > 1:
            return "NCCL batched send/recv chain pipelined over %d stages of tensors + broadcast" % self.parts
        return "NCCL send/recv chain + broadcast"

    def _record_per_parameter(self, user, scale):
        for i, param in enumerate(self.parameters):
            if user != 0:
                param.grad.data.add_(self.compressed_gradients[i][-1])
            if self.error_feedback:
                param.grad.data.add_(scale * param.error[user])
                decompressed_g = self.compressors[i].decompress(
                    self.compressors[i].compress(param.grad.data))
                param.error[user].data = param.grad.data - decompressed_g
            else:
                decompressed_g = self.compressors[i].decompress(
                    self.compressors[i].compress(param.grad.data))
            self.compressed_gradients[i].append(decompressed_g)

    def apply(self):
        if self.plan is None:
            for i, param in enumerate(self.parameters):
                param.grad.data = self.compressed_gradients[i][-1]
            for compressed in self.compressed_gradients:
                compressed.clear()
            return
        plan = self.plan
        last = self.args.num_users - 1
        if self.distributed:
            xch.ring_broadcast_last(plan.records, self.world)
        g = plan.decode(first_user=last, n_users=1, mean=False, out=plan.arena)
        self._set_grads_from(g)
