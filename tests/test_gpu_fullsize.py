"""GPU, BASELINE.json's full sizes (ResNet-50 gradient: 23 520 842 elements, 76 compressed tensors):
the CPU oracle over the WHOLE gradient (every chunk's code, u, level, every tensor's lb/ub; the C
port does a user-pass in about a second), the 8-user ps record/apply flow and the QSGD / TernGrad /
sign / top-k codecs at the same size against the oracle, two independent CUDA implementations
against each other, and the codec's algebraic invariants."""
import numpy as np
import pytest
import torch

import gq_b200
from gq_b200 import _lib
from gq_b200.quantizers.fused import FusedPlan
from oracle import gq_oracle as O
from util import codebook, make_args, resnet50_shapes

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def plans():
    shapes = resnet50_shapes()
    out = {}
    for name, algo in (("tc", _lib.ALGO_AUTO), ("exact", _lib.ALGO_EXACT)):
        a = make_args(num_users=2, hsq_algo=algo)
        out[name] = FusedPlan(gq_b200.NearestNeighborCompressor, shapes, a, torch.device(DEV), 2)
    gen = torch.Generator(device=DEV)
    gen.manual_seed(7)
    g = torch.randn(out["tc"].arena_elems, device=DEV, generator=gen) * 0.01
    # make it less uniform: per-tensor scales over 6 orders of magnitude, some exactly-zero tensors
    p = out["tc"]
    for i in range(len(p.sizes)):
        v = p.view(i, g)
        v.mul_(float(10.0 ** ((i % 7) - 4)))
        if i % 37 == 5:
            v.zero_()
    return out, g


def test_tcgen05_and_exact_kernels_agree_on_every_chunk(plans):
    pl, g = plans
    n = pl["tc"].groups[0].n_chunks
    uni = {id(pl[k].groups[0]): torch.rand(n, device=DEV, generator=torch.Generator(device=DEV).manual_seed(3))
           for k in pl}
    for k in pl:
        pl[k].encode(0, src=g, uniforms=uni)
    torch.cuda.synchronize()
    assert torch.equal(pl["tc"].records[0], pl["exact"].records[0])      # codes, levels, lb/ub, identity: all bytes
    assert torch.equal(pl["tc"].u_scratch[:n], pl["exact"].u_scratch[:n])


def test_oracle_agrees_on_every_chunk_of_the_whole_gradient(plans):
    """codes, u, l and lb/ub of all 1 468 652 chunks / 76 tensors against the CPU oracle."""
    pl, g = plans
    p = pl["tc"]
    grp = p.groups[0]
    n = grp.n_chunks
    uni = torch.rand(n, device=DEV, generator=torch.Generator(device=DEV).manual_seed(11))
    p.random = 1
    p.encode(1, src=g, uniforms={id(grp): uni})
    torch.cuda.synchronize()
    chunks = g[grp.arena_off:grp.arena_off + grp.n].view(-1, 16).cpu().numpy()
    oc, ou = O.hsq_search(chunks, codebook(16, 256))
    rec = p.records[1]
    assert np.array_equal(rec[grp.codes_off:grp.codes_off + n].cpu().numpy().astype(np.int32), oc)
    assert np.array_equal(p.u_scratch[:n].cpu().numpy(), ou)
    l = rec[grp.l_off:grp.l_off + n].cpu().numpy()
    lbub = rec[grp.lbub_off:grp.lbub_off + 8 * grp.n_seg].view(torch.float32).view(-1, 2).cpu().numpy()
    r = uni.cpu().numpy()
    starts = grp.seg_start_host
    for sg in range(grp.n_seg):
        a, b = starts[sg], starts[sg + 1]
        lb, ub, ol, used = O.psc_compress(ou[a:b], 6, True, r[a:b])
        assert lbub[sg, 0] == lb and lbub[sg, 1] == ub, sg
        assert np.array_equal(l[a:b].astype(np.int32), ol), sg


def test_eight_user_ps_step_equals_the_oracle_at_full_size():
    """PSQuantizer.record x 8 + apply() on the ResNet-50 shapes against oracle.ps_step
    (ps_quantizer.py:27-65), same external uniform stream; decode over U = 8 records at full size."""
    shapes = resnet50_shapes()
    sizes = [int(np.prod(s)) for s in shapes]
    U = 8
    a = make_args(num_users=U)
    params = [torch.nn.Parameter(torch.zeros(s, device=DEV)) for s in shapes]
    q = gq_b200.Quantizer(gq_b200.NearestNeighborCompressor, params, a)
    plan = q.plan
    gen = torch.Generator(device=DEV)
    gen.manual_seed(21)
    per_user = sum(n // 16 for n in sizes if n > 1000)
    rs = np.random.RandomState(5)
    draws = rs.random_sample(per_user * U).astype(np.float32)
    grads = []
    for u in range(U):
        flat = torch.randn(plan.arena_elems, device=DEV, generator=gen) * (0.01 * (u + 1))
        views = plan.views(flat)
        grads.append([v.cpu().numpy() for v in views])
        for p, v in zip(params, views):
            p.grad = v.clone()              # ordinary per-parameter tensors: the gather kernel runs
        parts, used = plan.split_uniform_stream(draws[u * per_user:(u + 1) * per_user])
        assert used == per_user
        q.record(u, epoch=1, uniforms=parts)
    q.apply()
    torch.cuda.synchronize()
    cb = codebook(16, 256)
    codecs = [O.HSQ(n, s, cb, 6, True) if n > 1000 else O.Identity() for n, s in zip(sizes, shapes)]
    ref = O.ps_step(codecs, grads, O.UniformStream(draws))
    for i, (p, r_) in enumerate(zip(params, ref)):
        got = p.grad.data.cpu().numpy()
        r_ = r_.reshape(got.shape)
        if got.size >= 256:
            assert np.array_equal(got, r_), i
        else:
            assert np.abs(got - r_).max() <= 1e-6 * max(np.abs(r_).max(), 1e-30), i


@pytest.mark.parametrize("quant,kw", [("qsgd", dict(c_dim=128, n_bit=2)), ("terngrad", dict(c_dim=0, n_bit=1)),
                                      ("sign", {}), ("topk", dict(cr=100))])
def test_elementwise_codecs_equal_the_oracle_at_full_size(quant, kw):
    """QSGD d=128 2-bit (first conv: dim 192), TernGrad, sign and top-k 1 % on the ResNet-50 shapes,
    two users through PSQuantizer.record/apply against the oracle (BASELINE configs 3 and 5)."""
    shapes = resnet50_shapes()
    sizes = [int(np.prod(s)) for s in shapes]
    U = 2
    a = make_args(num_users=U, **kw)
    Comp = {"qsgd": gq_b200.QSGDCompressor, "terngrad": gq_b200.QSGDCompressor, "sign": gq_b200.SignSGDCompressor,
            "topk": gq_b200.TopKSparsificationCompressor}[quant]
    params = [torch.nn.Parameter(torch.zeros(s, device=DEV)) for s in shapes]
    q = gq_b200.Quantizer(Comp, params, a)
    plan = q.plan
    gen = torch.Generator(device=DEV)
    gen.manual_seed(31)
    per_user = sum(n for n in sizes if n > 1000) if quant in ("qsgd", "terngrad") else 0
    draws = np.random.RandomState(6).random_sample(per_user * U).astype(np.float32)
    grads = []
    for u in range(U):
        flat = torch.randn(plan.arena_elems, device=DEV, generator=gen) * 0.01
        if quant == "qsgd":
            flat[1000:5000] = 0.0          # some all-zero chunks (the 0/0 -> INT32_MIN edge)
        views = plan.views(flat)
        grads.append([v.cpu().numpy() for v in views])
        for p, v in zip(params, views):
            p.grad = v.clone()
        parts, used = plan.split_uniform_stream(draws[u * per_user:(u + 1) * per_user])
        assert used == per_user
        q.record(u, epoch=1, uniforms=parts)
    q.apply()
    torch.cuda.synchronize()
    codecs = []
    for n, s in zip(sizes, shapes):
        if n <= 1000:
            codecs.append(O.Identity())
        elif quant in ("qsgd", "terngrad"):
            codecs.append(O.QSGD(n, s, a.c_dim, a.n_bit, True))
        elif quant == "sign":
            codecs.append(O.Sign(n, s))
        else:
            codecs.append(O.TopK(n, s, a.cr))
    ref = O.ps_step(codecs, grads, O.UniformStream(draws))
    for i, (p, r_) in enumerate(zip(params, ref)):
        got = p.grad.data.cpu().numpy()
        r_ = r_.reshape(got.shape)
        if got.size >= 256:
            assert np.array_equal(got, r_), (quant, i)
        else:
            assert np.abs(got - r_).max() <= 1e-6 * max(np.abs(r_).max(), 1e-30), (quant, i)


def test_levels_bounds_and_roundtrip_error(plans):
    pl, g = plans
    p = pl["tc"]
    grp = p.groups[0]
    p.encode(0, src=g)
    out = torch.zeros_like(g)
    p.decode(first_user=0, n_users=1, mean=False, out=out)
    torch.cuda.synchronize()
    l = p.records[0][grp.l_off:grp.l_off + grp.n_chunks]
    assert int(l.max()) <= 64 and int(l.min()) >= 0                       # 2^n + 1 levels
    lbub = p.records[0][grp.lbub_off:grp.lbub_off + 8 * grp.n_seg].view(torch.float32).view(-1, 2)
    u = p.u_scratch[:grp.n_chunks]
    starts = grp.seg_start_host
    for s in (0, 1, 17, 40, 75):
        seg_u = u[starts[s]:starts[s + 1]]
        assert lbub[s, 0].item() == seg_u.min().item() and lbub[s, 1].item() == seg_u.max().item()
    # identity tensors pass through untouched; all-zero tensors decode to zero
    for i, sz in enumerate(p.sizes):
        if sz <= 1000:
            assert torch.equal(p.view(i, out), p.view(i, g))
        elif i % 37 == 5:
            assert not p.view(i, out).any()
    # the decoded chunk is a scaled codeword: re-encoding it picks the same codeword (idempotence of
    # the search on its own output, up to the sign carried by the norm)
    x = g[grp.arena_off:grp.arena_off + grp.n].view(-1, 16)
    y = out[grp.arena_off:grp.arena_off + grp.n].view(-1, 16)
    # HSQ keeps the projection on one unit codeword: ||x - y||^2 = ||x||^2 - u^2 up to norm quantization
    nx2 = (x.double() ** 2).sum(1)
    err2 = ((x.double() - y.double()) ** 2).sum(1)
    ny = y.double().norm(dim=1)
    assert bool((err2 <= nx2 * (1 + 1e-6) + (ny - u.double().abs()) ** 2 + 1e-30).all())


def test_mean_of_identical_users_is_the_single_user_decode(plans):
    pl, g = plans
    p = pl["tc"]
    uni = {id(p.groups[0]): torch.rand(p.groups[0].n_chunks, device=DEV)}
    p.encode(0, src=g, uniforms=uni)
    p.encode(1, src=g, uniforms=uni)
    one = torch.zeros_like(g)      # zeros: the arena's alignment padding is never written
    two = torch.zeros_like(g)
    p.decode(first_user=0, n_users=1, mean=False, out=one)
    p.decode(first_user=0, n_users=2, mean=True, out=two)
    torch.cuda.synchronize()
    assert torch.equal(p.records[0], p.records[1])
    assert torch.equal(one, two)                                          # (d + d) / 2 == d exactly
    # linearity of the decoder in the users: decode-accumulate of user 1 onto user 0 == 2 * decode
    acc = one.clone()
    p.decode(first_user=1, n_users=1, mean=False, accumulate=True, out=acc)
    torch.cuda.synchronize()
    assert torch.equal(acc, one * 2)
