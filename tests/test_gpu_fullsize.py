"""GPU, BASELINE.json's full sizes (ResNet-50 gradient: 23 520 842 elements, 76 compressed tensors):
size-independent properties instead of a full CPU oracle pass -- two independent CUDA
implementations agree everywhere, the oracle agrees on a random sample of chunks, and the
codec's algebraic invariants hold."""
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


def test_oracle_agrees_on_a_random_sample_of_chunks(plans):
    pl, g = plans
    p = pl["tc"]
    grp = p.groups[0]
    p.encode(1, src=g)
    torch.cuda.synchronize()
    rs = np.random.RandomState(0)
    idx = np.unique(np.concatenate([rs.randint(0, grp.n_chunks, 60000), np.arange(0, 4096),
                                    np.arange(grp.n_chunks - 4096, grp.n_chunks)]))
    chunks = g[grp.arena_off:grp.arena_off + grp.n].view(-1, 16)[torch.from_numpy(idx).to(DEV)].cpu().numpy()
    oc, ou = O.hsq_search(chunks, codebook(16, 256))
    codes = p.records[1][grp.codes_off:grp.codes_off + grp.n_chunks].cpu().numpy()[idx]
    assert np.array_equal(codes.astype(np.int32), oc)
    assert np.array_equal(p.u_scratch[:grp.n_chunks].cpu().numpy()[idx], ou)


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
