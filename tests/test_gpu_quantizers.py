"""GPU parity tests of the ps / ring record-apply flow against the reference's own
outputs (golden fixtures) and the oracle, fused and per-parameter paths."""
import numpy as np
import pytest
import torch

import gq_b200
from oracle import gq_oracle as O
from util import gen_input, golden, golden_names, make_args, torch_uniform_stream

pytestmark = pytest.mark.gpu
DEV = "cuda"

COMP = {"hsq": "NearestNeighborCompressor", "qsgd": "QSGDCompressor", "sign": "SignSGDCompressor",
        "topk": "TopKSparsificationCompressor"}


def _run(g, fused):
    U, seed, iters = int(g["U"]), int(g["seed"]), int(g["iters"])
    shapes = [tuple(int(x) for x in g["shape%d" % i]) for i in range(int(g["n_tensors"]))]
    sizes = [int(np.prod(s)) for s in shapes]
    two_phase = bool(int(g["two_phase"])) if "two_phase" in g else False
    a = make_args(mode=str(g["mode"]), num_users=U, ef=bool(int(g["ef"])), two_phase=two_phase, c_dim=int(g["c_dim"]),
                  k_bit=int(g["k_bit"]), n_bit=int(g["n_bit"]), cr=int(g["cr"]), fused=fused)
    params = [torch.nn.Parameter(torch.zeros(s, device=DEV)) for s in shapes]
    q = gq_b200.Quantizer(getattr(gq_b200, COMP[str(g["quant"])]), params, a)
    assert (q.plan is not None) == fused
    stream = torch_uniform_stream(seed, int(g["n_draws"]))
    pos = 0
    for it in range(iters):
        for u in range(U):
            for i, (p, n, s) in enumerate(zip(params, sizes, shapes)):
                p.grad = torch.from_numpy(gen_input(seed * 1000 + it * 100 + u * 10 + i, n)).view(s).to(DEV)
            parts, used = q.plan.split_uniform_stream(stream[pos:])
            pos += used
            q.record(u, epoch=int(g["epoch"]), uniforms=parts)
        if two_phase:   # the second compression draws from the same stream, after every user's record
            parts, used = q.phase2_plan().split_uniform_stream(stream[pos:])
            pos += used
            q.apply(uniforms=parts)
        else:
            q.apply()
        for i, p in enumerate(params):
            ref = g["grad_it%d_t%d" % (it, i)]
            got = p.grad.data.cpu().numpy().reshape(-1)
            rel = np.abs(ref - got).max() / max(np.abs(ref).max(), 1e-30)
            assert rel <= 1e-5, (it, i, rel)                 # north_star bar
            if ref.size >= 256:
                assert np.array_equal(ref, got), (it, i, rel)  # what we actually hold
    assert pos == stream.size


@pytest.mark.parametrize("name", golden_names("ps_") + golden_names("ring_"))
def test_quantizer_fused_vs_reference_golden(name):
    _run(golden(name), fused=True)


def test_per_parameter_path_matches_fused():
    """fused=False runs the reference's per-parameter loop on the same kernels."""
    shapes = [(64, 128), (100,), (32, 64)]
    sizes = [int(np.prod(s)) for s in shapes]
    res = {}
    for fused in (True, False):
        a = make_args(num_users=3, fused=fused, random=False)
        params = [torch.nn.Parameter(torch.zeros(s, device=DEV)) for s in shapes]
        q = gq_b200.Quantizer(gq_b200.NearestNeighborCompressor, params, a)
        for u in range(3):
            for i, (p, n, s) in enumerate(zip(params, sizes, shapes)):
                p.grad = torch.from_numpy(gen_input(u * 10 + i, n)).view(s).to(DEV)
            q.record(u, epoch=0)
        q.apply()
        res[fused] = [p.grad.data.clone() for p in params]
    for a_, b_ in zip(res[True], res[False]):
        assert torch.equal(a_, b_)


def test_two_phase_and_identical_users():
    shapes = [(64, 128), (10,)]
    a = make_args(num_users=4, two_phase=True, random=False)
    params = [torch.nn.Parameter(torch.zeros(s, device=DEV)) for s in shapes]
    q = gq_b200.Quantizer(gq_b200.NearestNeighborCompressor, params, a)
    xs = [gen_input(5 + i, int(np.prod(s))).reshape(s) for i, s in enumerate(shapes)]
    for u in range(4):
        for p, x in zip(params, xs):
            p.grad = torch.from_numpy(x).to(DEV)
        q.record(u, epoch=0)
    q.apply()
    # identical users: mean == single-user decode; second phase compresses it again
    from util import codebook
    oc = O.HSQ(xs[0].size, shapes[0], codebook(16, 256), 6, False)
    once = oc.decompress(oc.compress(xs[0]))
    twice = oc.decompress(oc.compress(once))
    assert np.array_equal(params[0].grad.data.cpu().numpy(), twice)
    assert np.array_equal(params[1].grad.data.cpu().numpy(), xs[1])


def test_sgd_identity_quantizer():
    shapes = [(64, 128), (10,)]
    a = make_args(num_users=2)
    params = [torch.nn.Parameter(torch.zeros(s, device=DEV)) for s in shapes]
    q = gq_b200.Quantizer(gq_b200.IdenticalCompressor, params, a)
    xs = [[gen_input(u * 7 + i, int(np.prod(s))).reshape(s) for i, s in enumerate(shapes)] for u in range(2)]
    for u in range(2):
        for p, x in zip(params, xs[u]):
            p.grad = torch.from_numpy(x).to(DEV)
        q.record(u, epoch=0)
    q.apply()
    for i, p in enumerate(params):
        assert np.array_equal(p.grad.data.cpu().numpy(), (xs[0][i] + xs[1][i]) / np.float32(2))


@pytest.mark.parametrize("c_dim,n_bit,users", [(16, 6, 3), (16, 32, 2), (8, 6, 3), (32, 6, 2), (16, 6, 9)])
def test_identity_tensors_follow_every_hsq_path(c_dim, n_bit, users):
    """The small (identity) tensors ride inside the HSQ init / staged decode kernels when those
    run (d = 16, 6-bit norms, <= 8 users) and are launched on their own otherwise (fp32 norms:
    no init kernel; d = 8 / 32: other decode kernels; 9 users: no attachment): same results."""
    from util import codebook
    shapes = [(96, 128), (10,), (64, 64), (999,)]
    sizes = [int(np.prod(s)) for s in shapes]
    a = make_args(num_users=users, c_dim=c_dim, n_bit=n_bit, random=False)
    params = [torch.nn.Parameter(torch.zeros(s, device=DEV)) for s in shapes]
    q = gq_b200.Quantizer(gq_b200.NearestNeighborCompressor, params, a)
    assert q.plan is not None
    xs = [[gen_input(300 + 10 * u + i, n).reshape(s) for i, (n, s) in enumerate(zip(sizes, shapes))]
          for u in range(users)]
    for u in range(users):
        for p, x in zip(params, xs[u]):
            p.grad = torch.from_numpy(x).to(DEV)
        q.record(u, epoch=0)
    q.apply()
    codecs = [O.HSQ(n, s, codebook(O.chunk_dim(n, c_dim), 256), n_bit, False) if n > 1000 else O.Identity()
              for n, s in zip(sizes, shapes)]
    ref = O.ps_step(codecs, xs, O.UniformStream(np.zeros(0, np.float32)))
    for p, r_ in zip(params, ref):
        got = p.grad.data.cpu().numpy()
        r_ = r_.reshape(got.shape)
        if got.size >= 256:
            assert np.array_equal(got, r_)
        else:
            assert np.abs(got - r_).max() <= 1e-6 * max(np.abs(r_).max(), 1e-30)


def test_codebook_size_equal_to_chunk_dim_takes_the_per_parameter_path():
    """2 ** k_bit == chunk dim (c_dim 16, k_bit 4) asks for random orthogonal bases
    (nearest_neighbor_compressor.py:45): the quantizer must fall back to per-parameter compressors
    instead of failing in the fused plan (round-1 advisor finding)."""
    shapes = [(64, 128), (100,)]
    a = make_args(num_users=2, c_dim=16, k_bit=4, random=False)
    params = [torch.nn.Parameter(torch.zeros(s, device=DEV)) for s in shapes]
    q = gq_b200.Quantizer(gq_b200.NearestNeighborCompressor, params, a)
    assert q.plan is None
    for u in range(2):
        for i, (p, s) in enumerate(zip(params, shapes)):
            p.grad = torch.from_numpy(gen_input(u * 10 + i, int(np.prod(s)))).view(s).to(DEV)
        q.record(u, epoch=0)
    q.apply()
    assert all(torch.isfinite(p.grad).all() for p in params)


@pytest.mark.parametrize("algo", ["auto", "exact"])
def test_nan_gradient_makes_the_whole_tensor_nan(algo):
    """torch.min / torch.max propagate NaN in the reference (probabilistic_scalar_compressor.py:14-15):
    one NaN element turns the decoded tensor into NaN; the other tensors are untouched."""
    from gq_b200 import _lib
    shapes = [(64, 128), (32, 64)]
    a = make_args(num_users=1, random=False, hsq_algo={"auto": _lib.ALGO_AUTO, "exact": _lib.ALGO_EXACT}[algo])
    params = [torch.nn.Parameter(torch.zeros(s, device=DEV)) for s in shapes]
    q = gq_b200.Quantizer(gq_b200.NearestNeighborCompressor, params, a)
    xs = [gen_input(40 + i, int(np.prod(s))).reshape(s) for i, s in enumerate(shapes)]
    xs[0][3, 5] = np.nan
    for p, x in zip(params, xs):
        p.grad = torch.from_numpy(x).to(DEV)
    q.record(0, epoch=0)
    q.apply()
    assert torch.isnan(params[0].grad).all()
    assert torch.isfinite(params[1].grad).all()


@pytest.mark.parametrize("name", [n for n in golden_names("ring_") if "ef" not in n])
def test_ring_in_stages_equals_the_reference_golden(name, monkeypatch):
    """The pipelined ring cuts the plan into stages of tensors; hop by hop and stage by
    stage it must reproduce the reference's chain bit for bit (same Philox / uniform indices)."""
    monkeypatch.setenv("GQ_RING_PARTS", "3")
    g = golden(name)
    _run(g, fused=True)
